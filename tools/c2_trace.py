"""Kernel-level timeline (engine flag 32) of a blocking vsf_get_matches call on C2 (2000 x 2000): when the
expansion, distance and finish kernels start and end relative to the first of them.  Under gpurun:

    python tools/c2_trace.py
"""
import ctypes as C, sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import vision_slam_frontend_b200 as vsf
import synth
ctx = vsf.Context(device=0, max_features=2048, desc_bytes=32, window=1)
Q, T = synth.descriptor_pair(2000, 2000, seed=1)
R = float(np.float32(0.6))
for _ in (0,):
    ctx.set_engine(2, 32)
    for i in range(60):
        t0 = time.perf_counter(); m = ctx.get_matches(Q, T, R); t1 = time.perf_counter()
    buf = np.zeros((256, 8, 2), np.int64); got = C.c_int(0)
    ctx._check(ctx._L.vsf_debug_kernel_trace(ctx._h, buf.ctypes.data, 256, C.byref(got)))
    tr = buf[:got.value].astype(np.float64)
    for p in range(got.value - 3, got.value):
        t0k = min(tr[p, k, 0] for k in (0, 1, 2) if tr[p, k, 1] > 0)
        print({name: [round((tr[p, k, 0] - t0k) / 1e3, 2), round((tr[p, k, 1] - t0k) / 1e3, 2)] for k, name in ((0, "expand"), (1, "distance"), (2, "finish"), (4, "finish_after_wait")) if tr[p, k, 1] > 0 or tr[p,k,0] < 1e18})
    print("python call us", round((t1 - t0) * 1e6, 1))
    break
