"""Per-CTA timeline of knn2_tc_kernel (engine flag 16; needs a library built with the stamps:
`VSF_TC_TRACE=1 python -m vision_slam_frontend_b200.build --force`) on a window-matching shape: where the
time between kernel entry and exit goes (setup, wait for the predecessor, first tile, MMA
stream, tail).  Run under gpurun:

    python tools/tc_timeline.py [features] [window] [out.json]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NAMES = ["entry", "setup_done", "pred_done", "first_tma", "first_A", "first_tile_landed", "first_mma",
         "last_commit", "last_acc_done", "last_partial", "exit"]


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"tc_timeline_{n}_{W}.json")
    RATIO = float(np.float32(0.6))
    ctx = vsf.Context(device=0, max_features=max(n, 256), desc_bytes=32, window=W)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = ctx._L
    poses = 12
    seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
    base, fb = seq.data_ptr(), n * 32

    def step(t):
        qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
        nn = (C.c_int * W)(*([n] * W))
        rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
        assert rc == 0, L.vsf_last_error(ctx._h)

    ctx.set_engine(2, 16)
    for t in range(poses):
        step(t)
    sm = L.vsf_device_sm_count(ctx._h)
    buf = np.zeros((sm, 16), np.int64)
    got = C.c_int(0)
    ctx._check(L.vsf_debug_tc_trace(ctx._h, buf.ctypes.data, sm, C.byref(got)))
    ctx.set_engine(0, 0)
    tr = buf[: got.value]
    tr = tr[tr[:, 1] != 0]
    if len(tr) == 0:
        raise SystemExit("no stamps: rebuild with VSF_TC_TRACE=1 python -m vision_slam_frontend_b200.build --force")
    cyc = (tr[:, 11] - tr[:, 1]).astype(np.float64)
    ns = (tr[:, 15] - tr[:, 0]).astype(np.float64)
    ghz = float(np.median(cyc / np.maximum(ns, 1)))
    rel = (tr[:, 1:12] - tr[:, 1:2]) / ghz / 1e3          # us since the CTA's entry
    start_ns = tr[:, 0] - tr[:, 0].min()
    res = {
        "features": n, "window": W, "ctas": int(len(tr)), "sm_ghz": ghz,
        "kernel_span_us": float((tr[:, 15].max() - tr[:, 0].min()) / 1e3),
        "cta_entry_spread_us": float(start_ns.max() / 1e3),
        "segments_per_cta_max": int(tr[:, 14].max()),
        "mma_wait_for_queries_us_median": float(np.median(tr[:, 12]) / ghz / 1e3),
        "mma_wait_for_tiles_us_median": float(np.median(tr[:, 13]) / ghz / 1e3),
        "phases_us_since_cta_entry": {
            NAMES[i]: {"median": float(np.median(rel[:, i])), "max": float(rel[:, i].max()), "min": float(rel[:, i].min())}
            for i in range(len(NAMES))
        },
        "mma_stream_us_median": float(np.median(rel[:, 7] - rel[:, 6])),
        "tail_after_last_commit_us_median": float(np.median(rel[:, 10] - rel[:, 7])),
    }
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
