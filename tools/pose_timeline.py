"""Kernel-level timeline of a stream of window poses (engine flag 32): when the expansion,
distance, refine and compaction kernels of consecutive poses start and end on the device.  Run under gpurun:

    python tools/pose_timeline.py [features] [window] [out.json]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"pose_timeline_{n}_{W}.json")
    RATIO = float(np.float32(0.6))
    ctx = vsf.Context(device=0, max_features=max(n, 256), desc_bytes=32, window=W)
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = ctx._L
    poses = 48
    seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
    base, fb = seq.data_ptr(), n * 32

    def step(t):
        qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
        nn = (C.c_int * W)(*([n] * W))
        rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
        assert rc == 0, L.vsf_last_error(ctx._h)

    res = {}
    names = ["expand", "distance", "refine", "compact", "refine_after_wait"]
    for overlap in (0,):
        ctx.set_engine(2, 0)
        for t in range(8):
            step(t)
        ctx.synchronize()
        ctx.set_engine(2, 32)
        for t in range(poses):
            step(t)
        buf = np.zeros((256, 8, 2), np.int64)
        got = C.c_int(0)
        ctx._check(L.vsf_debug_kernel_trace(ctx._h, buf.ctypes.data, 256, C.byref(got)))
        tr = buf[: got.value].astype(np.float64)
        t0 = tr[24, 1, 0]
        rows = []
        for p in range(24, 29):
            rows.append({names[k]: [round((tr[p, k, 0] - t0) / 1e3, 2), round((tr[p, k, 1] - t0) / 1e3, 2)] for k in range(5)})
        period = float(np.median(np.diff(tr[8:, 1, 0])) / 1e3)
        res["overlap" if overlap else "serial"] = {
            "us_between_distance_kernel_starts_median": period,
            "distance_us_median": float(np.median(tr[8:, 1, 1] - tr[8:, 1, 0]) / 1e3),
            "refine_us_median": float(np.median(tr[8:, 2, 1] - tr[8:, 4, 0]) / 1e3),
            "poses_24_28_us_[start,end]": rows,
        }
        ctx.set_engine(0, 0)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
