"""Two / three / four kernels per pose in vsf_window_match_block_device (engine flags 0 / 256 / 512):
per-pose time of each and the kernel-level timeline (engine flag 32) of a few consecutive poses.  Run under gpurun:

    python tools/chain_probe.py [features] [window] [desc_bytes] [out.json]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    width = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    out_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "gpurun_out", f"chain_probe_{n}_{W}_{width}.json")
    RATIO = float(np.float32(0.6))
    ctx = vsf.Context(device=0, max_features=max(n, 256), desc_bytes=width, window=W)
    stream = torch.cuda.Stream(priority=-1)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    rb = ctx.row_bytes
    n_poses = max(2 * W + 2, int(1.5 * 126e6 / (n * rb)) + 1)
    seq = torch.empty((n_poses, n, rb), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, n_poses, max(1, n // 10), 7)
    base = seq.data_ptr()
    B = 512

    def timed(flags, reps=3):
        ctx.set_engine(2, flags)
        ctx.window_match_block_device(base, n, n_poses, 0, B, RATIO)
        ctx.synchronize()
        best = 1e9
        for r in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s in range(4):
                ctx.window_match_block_device(base, n, n_poses, (1 + s) * B, B, RATIO)
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / (4 * B))
        return best

    res = {"features": n, "window": W, "desc_bytes": width}
    res["us_per_pose_two_kernels"] = timed(0)
    res["us_per_pose_three_kernels"] = timed(256)
    res["us_per_pose_four_kernels"] = timed(512)
    res["us_per_pose_two_kernels_no_early_start"] = timed(1024)
    res["us_per_pose_two_kernels_again"] = timed(0)
    counts = ctx.fetch_window(W, with_matches=False)
    res["survivors_last_pose"] = [int(c) for c in counts]

    # timeline of one chained block
    names = ["expand", "distance", "refine", "compact", "refine_after_wait"]
    for label, flags in (("two_kernels", 32), ("two_kernels_no_early_start", 32 | 1024), ("four_kernels", 32 | 512)):
        ctx.set_engine(2, flags)
        ctx.window_match_block_device(base, n, n_poses, 3, 40, RATIO)
        buf = np.zeros((256, 8, 2), np.int64)
        got = C.c_int(0)
        ctx._check(ctx._L.vsf_debug_kernel_trace(ctx._h, buf.ctypes.data, 256, C.byref(got)))
        tr = buf[: got.value].astype(np.float64)
        t0 = tr[20, 1, 0]
        rows = []
        for p in range(20, 24):
            rows.append({names[k]: [round((tr[p, k, 0] - t0) / 1e3, 2), round((tr[p, k, 1] - t0) / 1e3, 2)] for k in range(5)})
        extra = {}
        res["timeline_" + label] = {
            **extra,
            "us_between_distance_kernel_starts_median": float(np.median(np.diff(tr[4:, 1, 0])) / 1e3),
            "distance_us_median": float(np.median(tr[4:, 1, 1] - tr[4:, 1, 0]) / 1e3),
            "gap_distance_end_to_next_start_median": float(np.median(tr[5:, 1, 0] - tr[4:-1, 1, 1]) / 1e3),
            "poses_20_23_us_[start,end]": rows,
        }
    ctx.set_engine(0, 0)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
