"""Time the window kernel for one (engine, flags, split) setting: env N, W, POSES, ENGINE, FLAGS, SPLIT."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vision_slam_frontend_b200 as vsf

n = int(os.environ.get("N", 5000)); W = int(os.environ.get("W", 10)); poses = int(os.environ.get("POSES", 64))
engine = int(os.environ.get("ENGINE", 2)); flags = int(os.environ.get("FLAGS", 0)); split = int(os.environ.get("SPLIT", 0))
ctx = vsf.Context(device=0, max_features=n, desc_bytes=32, window=W)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
ctx.set_engine(engine, flags); ctx.set_tuning(-1, split, 0, -1)
seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
base, fb, L = seq.data_ptr(), n * 32, ctx._L
RATIO = float(np.float32(0.6))


def step(t):
    qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
    nn = (C.c_int * W)(*([n] * W))
    assert L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO) == 0


for t in range(5):
    step(t)
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for t in range(poses):
        step(t)
    e1.record(s)
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / poses)
print(json.dumps(dict(n=n, W=W, engine=engine, flags=flags, split=split, ms=round(best, 5),
                      gcmp_s=round(W * n * n / (best * 1e-3) / 1e9, 1))), flush=True)
