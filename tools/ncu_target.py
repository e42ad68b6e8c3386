"""Short target for ncu: a few launches of the window kernel on the C4 shape."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import vision_slam_frontend_b200 as vsf

n = int(os.environ.get("N", 5000)); W = int(os.environ.get("W", 10))
mode = int(os.environ.get("MODE", 0)); R = int(os.environ.get("QPT", 0)); split = int(os.environ.get("SPLIT", 0))
launches = int(os.environ.get("LAUNCHES", 6)); engine = int(os.environ.get("ENGINE", 0))
width = int(os.environ.get("WIDTH", 32))
ctx = vsf.Context(device=0, max_features=n, desc_bytes=width, window=W)
rb = ctx.row_bytes
s = torch.cuda.Stream(); torch.cuda.set_stream(s); ctx.set_stream(s.cuda_stream)
ctx.set_tuning(mode if engine <= 1 else -1, split, R, -1)
ctx.set_engine(engine, int(os.environ.get("FLAGS", 0)))
seq = torch.empty((launches + W, n, rb), dtype=torch.uint8, device="cuda")
ctx.synth_sequence_device(seq.data_ptr(), n, 0, launches + W, max(1, n // 10), 7)
base, fb = seq.data_ptr(), n * rb
for t in range(launches):
    qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
    nn = (C.c_int * W)(*([n] * W))
    assert ctx._L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, float(np.float32(0.6))) == 0
torch.cuda.synchronize()
print("done")
