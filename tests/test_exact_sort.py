"""Host side of sort_mode 1 (csrc/exact_sort.h): only the part of libstdc++'s introsort that can
reach the first int(n * best_percent) positions is carried out.  The kept prefix must be
exactly the one a full std::sort by distance leaves (src/slam_frontend.cc:289-291), order inside
equal-distance groups included; the checker is libstdc++'s own std::sort on 16-byte DMatch
records (oracle/stdsort_oracle.cc).  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import vision_slam_frontend_b200 as vsf
from oracle import restate

SHIFT = 22


def kept_prefix(dist: np.ndarray, keep: int) -> np.ndarray:
    keys = (dist.astype(np.uint32) << SHIFT) | np.arange(len(dist), dtype=np.uint32)
    keys = np.ascontiguousarray(keys)
    rc = vsf.load_library().vsf_debug_sort_prefix(keys.ctypes.data, len(keys), keep)
    assert rc == 0
    # the whole array is still a permutation of the input
    assert np.array_equal(np.sort(keys & np.uint32((1 << SHIFT) - 1)), np.arange(len(dist), dtype=np.uint32))
    return (keys[:keep] & np.uint32((1 << SHIFT) - 1)).astype(np.int32)


def oracle_order(dist: np.ndarray) -> np.ndarray:
    m = np.zeros(len(dist), dtype=vsf.DMATCH_DTYPE)
    m["queryIdx"] = np.arange(len(dist))
    m["distance"] = dist.astype(np.float32)
    return restate.sort_order_stdsort(m)


@pytest.mark.parametrize("best_percent", [0.3, 1.0, 0.0, 0.05, 0.77])
def test_prefix_equals_full_std_sort(best_percent):
    rng = np.random.default_rng(int(best_percent * 1000) + 3)
    sizes = list(range(0, 70)) + [100, 255, 256, 257, 500, 1000, 2500, 4500, 5000, 12000]
    for n in sizes:
        for spread in (1, 2, 7, 40, 257):
            dist = rng.integers(0, spread, size=n)
            keep = restate.num_good_matches(n, np.float32(best_percent))
            np.testing.assert_array_equal(kept_prefix(dist, keep), oracle_order(dist)[:keep])


def test_prefix_on_ordered_and_adversarial_inputs():
    rng = np.random.default_rng(5)
    for n in (17, 64, 1000, 4097):
        keep = restate.num_good_matches(n, restate.BEST_PERCENT)
        cases = {
            "sorted": np.arange(n) * 200 // max(n, 1),
            "reversed": (n - np.arange(n)) * 200 // max(n, 1),
            "all_equal": np.full(n, 9),
            "two_values": (np.arange(n) & 1) * 3,
            "organ_pipe": np.minimum(np.arange(n), n - 1 - np.arange(n)) % 256,
            "c4_like": 5 + rng.integers(0, 16, n) + rng.integers(0, 16, n) + rng.integers(0, 16, n),
        }
        for name, dist in cases.items():
            np.testing.assert_array_equal(kept_prefix(dist, keep), oracle_order(dist)[:keep], err_msg=f"{name} n={n}")


def test_prefix_rejects_bad_arguments():
    L = vsf.load_library()
    keys = np.zeros(4, np.uint32)
    assert L.vsf_debug_sort_prefix(keys.ctypes.data, 4, 5) != 0
    assert L.vsf_debug_sort_prefix(None, 4, 1) != 0
    assert L.vsf_debug_sort_prefix(None, 0, 0) == 0
