"""Where the device time of a host-API window step goes: per-kernel event times with and
without the mapped-host mirror, and the H2D copy."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vision_slam_frontend_b200 as vsf

n, W = 5000, 10
ctx = vsf.Context(device=0, max_features=n, desc_bytes=32, window=W)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
L = ctx._L
P = 64
seq = torch.empty((P + W, n, 32), dtype=torch.uint8, device="cuda")
ctx.synth_sequence_device(seq.data_ptr(), n, 0, P + W, n // 10, 7)
host = torch.empty((P + W, n, 32), dtype=torch.uint8).pin_memory(); host.copy_(seq); torch.cuda.synchronize()
hp = host.numpy()
RATIO = float(np.float32(0.6))
for p in range(W):
    ctx.window_push(p, hp[p])
ctx.set_profile(True)
fids = np.zeros(W, np.uint64); counts = np.zeros(W, np.int32)
out = np.zeros((W, n), dtype=vsf.DMATCH_DTYPE); nf = C.c_int(0)
kt = np.zeros(4); m = 0
for t in range(P):
    D = hp[W + t]
    assert L.vsf_window_match(ctx._h, D.ctypes.data, n, 32, RATIO, fids.ctypes.data, counts.ctypes.data, out.ctypes.data, n, C.byref(nf)) == 0
    if t >= 8:
        kt += np.array(ctx.last_kernel_times()); m += 1
    L.vsf_window_commit(ctx._h, W + t, n)
print("mirrored  [expand, main, refine, compact] us:", (kt / m * 1e3).round(2))
base, fb = seq.data_ptr(), n * 32
kt = np.zeros(4); m = 0
for t in range(P):
    qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
    nn = (C.c_int * W)(*([n] * W))
    assert L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO) == 0
    if t >= 8:
        kt += np.array(ctx.last_kernel_times()); m += 1
print("device    [expand, main, refine, compact] us:", (kt / m * 1e3).round(2))
ctx.set_profile(False)
# H2D copy of one frame, event-timed
d = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for t in range(40):
    e0.record(s); d.copy_(host[t], non_blocking=True); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("H2D 160 KB us:", round(1e3 * float(np.median(ts)), 2))
# pipelined loop: device time vs wall time
for mode in (1, 0):
    ctx.window_clear()
    for p in range(W):
        ctx.window_push(p, hp[p])
    fm = np.zeros((W, n), dtype=vsf.FEATURE_MATCH_DTYPE); fid = C.c_uint64(0)
    def collect():
        assert L.vsf_window_collect(ctx._h, C.byref(fid), fids.ctypes.data, counts.ctypes.data, fm.ctypes.data, n, C.byref(nf)) == 0
    for rep in range(2):
        torch.cuda.synchronize(); e0.record(s); t0 = time.perf_counter(); sub = 0.0
        for t in range(P):
            D = hp[W + t]
            a = time.perf_counter()
            assert L.vsf_window_submit(ctx._h, W + t + rep * P, D.ctypes.data, n, 32, RATIO, 0.3, mode, 1) == 0
            sub += time.perf_counter() - a
            if L.vsf_window_in_flight(ctx._h) > 2:
                collect()
        while L.vsf_window_in_flight(ctx._h):
            collect()
        e1.record(s); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print("pipelined mode", mode, "wall us/step", round(1e6 * wall / P, 1), "device-span us/step", round(1e3 * e0.elapsed_time(e1) / P, 1),
          "submit us/step", round(1e6 * sub / P, 1))
