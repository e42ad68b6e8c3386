"""vision_slam_frontend_b200 — B200-native (sm_100a) matching / stereo /
triangulation hot path of ut-amrl/vision_slam_frontend.

The product is `libvsf_cuda.so` (hand-written CUDA behind the C ABI declared in
`include/vsf.h`) plus the C++ `slam::Frontend` mirror in `csrc/frontend/`.
This Python package is only the ctypes view of that ABI used by the tests and
by `bench.py`, and the pose-range sharding / match-list gather used for the
multi-GPU runs.  There is no CPU fallback: importing works anywhere, but every
compute call raises if the CUDA library or a device is missing.
"""
from .capi import (VsfError, Context, DMATCH_DTYPE, KEYPOINT_DTYPE,  # noqa: F401
                   FEATURE_MATCH_DTYPE, PIPELINE_DEPTH, load_library, library_path)

__version__ = "0.1.0"
