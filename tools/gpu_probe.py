"""Integer-pipe probe + kernel-variant sweep (run on the GPU box under gpurun)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ctypes as C

import numpy as np
import torch

import vision_slam_frontend_b200 as vsf

out = {}
n, W = 5000, 10
ctx = vsf.Context(device=0, max_features=20000, desc_bytes=32, window=W)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
kinds = {0: "popc", 1: "lop3", 2: "popc+lop3", 3: "imad", 4: "vimnmx", 5: "popc+2lop3", 6: "iadd+xor"}
out["probe"] = {name: ctx.probe_pipe(k, 8192) for k, name in kinds.items()}
out["sm_count"] = ctx.sm_count
print(json.dumps(out["probe"], indent=1), flush=True)

RATIO = float(np.float32(0.6))


def time_window(n, W, poses=64, reps=3):
    seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
    base = seq.data_ptr()
    fb = n * 32
    L = ctx._L

    def step(t):
        qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
        nn = (C.c_int * W)(*([n] * W))
        rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
        assert rc == 0, L.vsf_last_error(ctx._h)
    for t in range(5):
        step(t)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(poses):
            step(t)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / poses)
    return best


res = []


def run(nn_, WW, poses, mode, R, split, var):
    ctx.set_tuning(mode, split, R, var)
    ms = time_window(nn_, WW, poses)
    r = dict(n=nn_, W=WW, mode=mode, R=R, split=split, var=var, ms=ms, gcmp_s=WW * nn_ * nn_ / (ms * 1e-3) / 1e9)
    res.append(r)
    print(r, flush=True)


for (nn_, WW, poses) in [(5000, 10, 64), (20000, 10, 4)]:
    for mode in (0, 2, 3):
        for var in (0, 1, 2, 3):
            for split in ((0, 1, 2, 4, 8) if nn_ == 5000 else (0, 1, 2)):
                run(nn_, WW, poses, mode, 1, split, var)
    for R in (2, 4):
        for var in (0, 1):
            run(nn_, WW, poses, 2, R, 0, var)
            run(nn_, WW, poses, 0, R, 0, var)
for mode in (0, 2):
    for var in (0, 1, 2):
        for split in (0, 1, 2, 4, 7):
            run(2000, 1, 200, mode, 1, split, var)
out["sweep"] = res
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_sweep.json"), "w"), indent=1)
