"""Model of the device-side emulation of libstdc++'s std::sort (csrc/exact_sort_kernel.cu), in
numpy, checked against the host implementation (csrc/exact_sort.h through vsf_debug_sort_prefix,
which itself is checked against the real std::sort in tests/test_exact_sort.py).

The point of the model is the PARALLEL formulation of libstdc++'s unguarded Hoare partition:

    while (true) { while (*first < pivot) ++first;  --last;  while (pivot < *last) --last;
                   if (!(first < last)) return first;  iter_swap(first, last);  ++first; }

pairs the k-th element from the left that is not < pivot (position L[k]) with the k-th element
from the right that is not > pivot (position R[k]) and swaps them while L[k] < R[k]; with k* the
first k that fails, the returned cut is min(L[k*], R[k*-1]) (a missing entry counts as +inf).
Both lists come from the ORIGINAL contents of the range, so all swaps of a partition step are
independent: ballots + prefix sums instead of a serial two-pointer walk.
"""
import numpy as np

SHIFT = 22
THRESH = 16


def _d(keys):
    return keys >> SHIFT


def partition_pivot(a, f, l):
    """std::__unguarded_partition_pivot on a[f:l) in place, parallel formulation; returns cut."""
    mid = f + (l - f) // 2
    ia, ib, ic = f + 1, mid, l - 1
    da, db, dc = a[ia] >> SHIFT, a[ib] >> SHIFT, a[ic] >> SHIFT
    # std::__move_median_to_first(result = f, a, b, c)
    if da < db:
        pick = ib if db < dc else (ic if da < dc else ia)
    else:
        pick = ia if da < dc else (ic if db < dc else ib)
    a[f], a[pick] = a[pick], a[f]
    piv = a[f] >> SHIFT
    seg = a[f + 1:l] >> SHIFT
    pos = np.arange(f + 1, l)
    L = pos[seg >= piv]                 # not (elem < pivot), left to right
    R = pos[seg <= piv][::-1]           # not (pivot < elem), right to left
    m = min(len(L), len(R))
    ok = L[:m] < R[:m]
    ks = int(ok.sum())                  # monotone: True ... True False ... False
    assert ok[:ks].all() and not ok[ks:].any()
    if ks:
        li, ri = L[:ks], R[:ks]
        tmp = a[li].copy()
        a[li] = a[ri]
        a[ri] = tmp
    big = 1 << 40
    cut = min(int(L[ks]) if ks < len(L) else big, int(R[ks - 1]) if ks >= 1 else big)
    assert f < cut <= l
    return cut


def partition_pivot_ranked(a, f, l):
    """The same step without the L list and without a search for k*: an element's own rank and
    the number of candidates on its other side decide whether it is swapped (what
    csrc/sort_kernel.cu does).  L element p of rank k (from the left) is swapped iff at least
    k + 1 R elements lie to its right (R[k] > p); R element r of rank k (from the right) iff at
    least k + 1 L elements lie to its left (L[k] < r).  The cut is the smallest position among
    the L elements that are NOT swapped and the R elements that ARE."""
    mid = f + (l - f) // 2
    ia, ib, ic = f + 1, mid, l - 1
    da, db, dc = a[ia] >> SHIFT, a[ib] >> SHIFT, a[ic] >> SHIFT
    if da < db:
        pick = ib if db < dc else (ic if da < dc else ia)
    else:
        pick = ia if da < dc else (ic if db < dc else ib)
    a[f], a[pick] = a[pick], a[f]
    piv = a[f] >> SHIFT
    seg = a[f + 1:l] >> SHIFT
    pos = np.arange(f + 1, l)
    isL, isR = seg >= piv, seg <= piv
    prefL = np.concatenate([[0], np.cumsum(isL)])          # L candidates in [f + 1, f + 1 + i)
    prefR = np.concatenate([[0], np.cumsum(isR)])
    totR = prefR[-1]
    i = np.arange(len(seg))
    rankL = prefL[i]                                       # rank from the left of an L element
    r_after = totR - prefR[i + 1]                          # R candidates strictly to the right
    rankR = r_after                                        # rank from the right of an R element
    l_before = prefL[i]                                    # L candidates strictly to the left
    swapL = isL & (r_after >= rankL + 1)
    swapR = isR & (l_before >= rankR + 1)
    cand = (isL & ~swapL) | swapR
    assert cand.any()
    cut = int(pos[cand].min())
    Rpos = np.zeros(len(seg) + 1, np.int64)
    Rpos[rankR[isR]] = pos[isR]                            # the R list, by rank from the right
    li = pos[swapL]
    ri = Rpos[rankL[swapL]]
    tmp = a[li].copy()
    a[li] = a[ri]
    a[ri] = tmp
    return cut


def heapsort(a, f, l):
    """std::__partial_sort(first, last, last) = make_heap + sort_heap, libstdc++'s sift rules."""
    v = a[f:l].copy()
    n = len(v)

    def less(x, y):
        return (x >> SHIFT) < (y >> SHIFT)

    def adjust(hole, length, value):
        top = hole
        child = hole
        while child < (length - 1) // 2:
            child = 2 * (child + 1)
            if less(v[child], v[child - 1]):
                child -= 1
            v[hole] = v[child]
            hole = child
        if (length & 1) == 0 and child == (length - 2) // 2:
            child = 2 * (child + 1)
            v[hole] = v[child - 1]
            hole = child - 1
        parent = (hole - 1) // 2
        while hole > top and less(v[parent], value):
            v[hole] = v[parent]
            hole = parent
            parent = (hole - 1) // 2
        v[hole] = value

    if n >= 2:
        parent = (n - 2) // 2
        while True:
            adjust(parent, n, v[parent])
            if parent == 0:
                break
            parent -= 1
        last = n
        while last > 1:
            last -= 1
            value = v[last]
            v[last] = v[0]
            adjust(0, last, value)
    a[f:l] = v


def sort_prefix_model(dist, keep, depth_limit=None, partition=None):
    partition = partition or partition_pivot
    n = len(dist)
    a = (np.asarray(dist, np.int64) << SHIFT) | np.arange(n, dtype=np.int64)
    if n == 0 or keep <= 0:
        return a[:0] & ((1 << SHIFT) - 1)
    keep = min(keep, n)
    if depth_limit is None:
        depth_limit = 2 * (int(n).bit_length() - 1)
    sorted_end = n
    segs = [(0, n, depth_limit)]
    levels = 0
    while segs:
        nxt = []
        for f, l, d in segs:                       # independent: one warp / CTA slice each
            if l - f <= THRESH:
                continue
            if d == 0:
                heapsort(a, f, l)
                continue
            cut = partition(a, f, l)
            if cut < keep:
                nxt.append((cut, l, d - 1))
            else:
                sorted_end = min(sorted_end, cut)
            nxt.append((f, cut, d - 1))
        segs = nxt
        levels += 1
    head = a[:sorted_end]
    order = np.argsort(head >> SHIFT, kind="stable")   # __final_insertion_sort == a stable sort
    return (head[order][:keep] & ((1 << SHIFT) - 1)).astype(np.int32), levels


if __name__ == "__main__":
    import ctypes
    import os
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import vision_slam_frontend_b200 as vsf
    lib = vsf.load_library()

    def host(dist, keep, depth=-1):
        keys = np.ascontiguousarray((dist.astype(np.uint32) << SHIFT) | np.arange(len(dist), dtype=np.uint32))
        if depth < 0:
            assert lib.vsf_debug_sort_prefix(keys.ctypes.data, len(keys), keep) == 0
        else:
            assert lib.vsf_debug_sort_prefix_depth(keys.ctypes.data, len(keys), keep, depth) == 0
        return (keys[:keep] & np.uint32((1 << SHIFT) - 1)).astype(np.int32)

    rng = np.random.default_rng(1)
    cases = 0
    maxlev = 0
    for n in list(range(0, 80)) + [100, 255, 256, 257, 500, 1000, 2500, 4500, 5000, 12000]:
        for spread in (1, 2, 7, 40, 257):
            for bp in (0.3, 1.0, 0.05):
                dist = rng.integers(0, spread, size=n)
                keep = int(np.float32(n) * np.float32(bp))
                got, lev = sort_prefix_model(dist, keep) if n and keep else (np.zeros(0, np.int32), 0)
                exp = host(dist, keep)
                assert np.array_equal(got, exp), (n, spread, bp)
                if n and keep:
                    got2, _ = sort_prefix_model(dist, keep, partition=partition_pivot_ranked)
                    assert np.array_equal(got2, exp), ("ranked", n, spread, bp)
                maxlev = max(maxlev, lev)
                cases += 1
    print("model == host exact sort on", cases, "cases; max levels", maxlev)
    if hasattr(lib, "vsf_debug_sort_prefix_depth"):
        for n in (40, 100, 1000, 4500):
            for depth in (0, 1, 2, 3):
                for spread in (2, 40, 300):
                    dist = rng.integers(0, spread, size=n)
                    keep = int(0.3 * n)
                    got, _ = sort_prefix_model(dist, keep, depth)
                    assert np.array_equal(got, host(dist, keep, depth)), (n, depth, spread)
        print("depth-limited (heapsort) cases ok")
