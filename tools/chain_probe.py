"""Groups of poses as one batch / pose by pose / four kernels per pose in vsf_window_match_block_device (engine flags 0 / 256 / 512):
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
    from vision_slam_frontend_b200 import capi
    for g in (4, 1, 4, 1, 4):
        ctx.set_option(capi.OPT_POSE_GROUP, g)
        res.setdefault("us_per_pose_group", {})[f"{g}" if f"{g}" not in res.get("us_per_pose_group", {}) else f"{g}_again"] = timed(0)
    res["us_per_pose_two_kernels"] = timed(0)
    res["us_per_pose_three_kernels"] = timed(256)
    res["us_per_pose_four_kernels"] = timed(512)
    res["us_per_pose_two_kernels_again"] = timed(0)
    counts = ctx.fetch_window(W, with_matches=False)
    res["survivors_last_pose"] = [int(c) for c in counts]

    # kernel-level timeline of a block of poses (one record per kernel sequence launched: with pose
    # groups a pose has one record for its distance kernel and one for its finish kernel)
    for label, flags, group in (("group_4", 32, 4), ("group_1", 32, 1), ("four_kernels", 32 | 512, 4)):
        ctx.set_option(capi.OPT_POSE_GROUP, group)
        ctx.set_engine(2, flags)
        ctx.window_match_block_device(base, n, n_poses, 3, 40, RATIO)
        buf = np.zeros((256, 8, 2), np.int64)
        got = C.c_int(0)
        ctx._check(ctx._L.vsf_debug_kernel_trace(ctx._h, buf.ctypes.data, 256, C.byref(got)))
        tr = buf[: got.value].astype(np.float64)
        recs = []
        # (a group of poses is one record: 40 poses = 10 records at group 4)
        lo = 3 if group > 1 and not (flags & 512) else 16
        t0 = None
        for p in range(lo, min(lo + 16, got.value)):
            r = {}
            for k, name in ((1, "distance"), (2, "finish")):
                if tr[p, k, 1] > 0:   # the kernel ran in this record
                    if t0 is None:
                        t0 = tr[p, k, 0]
                    r[name] = [round((tr[p, k, 0] - t0) / 1e3, 2), round((tr[p, k, 1] - t0) / 1e3, 2)]
            recs.append(r)
        res["timeline_" + label] = recs
    ctx.set_engine(0, 0)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
