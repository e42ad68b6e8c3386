"""No GPU needed: libvsf_nccl.so exports what include/vsf_nccl.h declares, and the C++ synthetic
stereo source is deterministic and counter-based (any pose on any rank)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_library_exports_every_declared_symbol():
    from vision_slam_frontend_b200 import nccl
    src = open(os.path.join(ROOT, "include", "vsf_nccl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = sorted(set(re.findall(r"\b(vsf_[a-z0-9_]+)\s*\(", src)))
    lib = nccl.load_nccl_library()
    assert len(decl) >= 6
    for name in decl:
        assert hasattr(lib, name), name
    assert sorted(nccl.EXPORTED_SYMBOLS) == decl


def test_synthetic_source_is_counter_based_and_deterministic():
    from vision_slam_frontend_b200.frontend import synth_frame, synthetic_rig
    a = synth_frame(400, 61, 9, 17)
    b = synth_frame(400, 61, 9, 17)
    c = synth_frame(400, 61, 9, 18)
    for x, y in zip(a[:6], b[:6]):
        np.testing.assert_array_equal(x, y)
    assert a[6] == b[6] and c[6] > a[6]
    assert (a[1] != c[1]).any() and a[1].shape == (400, 61) and a[1].dtype == np.uint8
    # consecutive poses share landmarks: a tenth of the features is new per pose
    rig = synthetic_rig()
    assert abs(np.abs(rig["fundamental"]).max()) > 0
    # the right image holds a permutation of the left features' descriptors (up to bit flips and outliers)
    dl, dr = a[1].astype(np.int16), a[3].astype(np.int16)
    pop = np.unpackbits((a[1][:, None, :8] ^ a[3][None, :, :8]), axis=2).sum(2)
    assert (pop.min(1) <= 6).mean() > 0.8


def test_pose_shard_halo():
    from vision_slam_frontend_b200 import sharding
    assert sharding.observe_halo_start(0, 10) == 0
    assert sharding.observe_halo_start(5, 10) == 0
    assert sharding.observe_halo_start(40, 10) == 29
