"""Kernel-level timeline (engine flag 32) of the host-buffer sequence path: vsf_window_run_sequence
over a few dozen frames, one record per launched batch.  Under gpurun:

    python tools/seq_probe.py [features] [window] [desc_bytes] [sort_mode] [lag] [pose_group]
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    width = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    sort_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    lag = int(sys.argv[5]) if len(sys.argv) > 5 else 12
    RATIO = float(np.float32(0.6))
    ctx = vsf.Context(device=0, max_features=n, desc_bytes=width, window=W)
    rb = ctx.row_bytes
    P = 1024 + W
    seq = torch.empty((P, n, rb), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, P, max(1, n // 10), 7)
    host = torch.empty((P, n, rb), dtype=torch.uint8).pin_memory()
    host.copy_(seq)
    torch.cuda.synchronize()
    hp = host.numpy()
    for p in range(W):
        ctx.window_push(p, hp[p])
    ring = 4
    out = np.zeros((ring, W, n), dtype=vsf.FEATURE_MATCH_DTYPE)
    cnts = np.zeros((ring, W), np.int32)
    flags = int(os.environ.get("VSF_ENGINE_FLAGS", "0"))
    ctx.set_engine(2, flags)
    if len(sys.argv) > 6:
        from vision_slam_frontend_b200 import capi
        ctx.set_option(capi.OPT_POSE_GROUP, int(sys.argv[6]))
    ctx.window_run_sequence(hp, W, 256, RATIO, 0.3, sort_mode, lag, out, cnts)   # warm-up
    t0 = time.perf_counter()
    ctx.window_run_sequence(hp, W + 256, 512, RATIO, 0.3, sort_mode, lag, out, cnts)
    wall = time.perf_counter() - t0
    res = {"features": n, "window": W, "desc_bytes": width, "sort_mode": sort_mode, "lag": lag,
           "us_per_frame": round(1e6 * wall / 512, 2)}
    ctx.set_engine(2, 32 | flags)
    ctx.window_run_sequence(hp, W + 768, 96, RATIO, 0.3, sort_mode, lag, out, cnts)
    buf = np.zeros((256, 8, 2), np.int64)
    got = C.c_int(0)
    ctx._check(ctx._L.vsf_debug_kernel_trace(ctx._h, buf.ctypes.data, 256, C.byref(got)))
    tr = buf[: got.value].astype(np.float64)
    recs, t0 = [], None
    for p in range(4, min(16, got.value)):
        r = {}
        for k, name in ((1, "distance"), (2, "finish")):
            if tr[p, k, 1] > 0:
                if t0 is None:
                    t0 = tr[p, k, 0]
                r[name] = [round((tr[p, k, 0] - t0) / 1e3, 2), round((tr[p, k, 1] - t0) / 1e3, 2)]
        recs.append(r)
    res["timeline"] = recs
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
