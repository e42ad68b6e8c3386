"""Per-pose time of vsf_window_match_block_device for pose groups 1, 2, 4 (tensor engine) - run once
per VSF_FIN_QPB setting to tune the finish kernel's queries per CTA.  Under gpurun:

    VSF_FIN_QPB=128 python tools/fin_probe.py [features] [window] [desc_bytes]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf
    from vision_slam_frontend_b200 import capi

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    width = int(sys.argv[3]) if len(sys.argv) > 3 else 32
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
    B = 256
    ctx.set_engine(2, int(os.environ.get("VSF_ENGINE_FLAGS", "0")))

    def timed(reps=3):
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
        return round(best, 2)

    res = {"qpb": os.environ.get("VSF_FIN_QPB", "auto"), "flags": os.environ.get("VSF_ENGINE_FLAGS", "0"), "features": n, "window": W, "desc_bytes": width, "us_per_pose": {}}
    for g in (1, 2, 4, 1, 2, 4):
        ctx.set_option(capi.OPT_POSE_GROUP, g)
        res["us_per_pose"].setdefault(str(g), []).append(timed())
    print(json.dumps(res))
    ctx.close()


if __name__ == "__main__":
    main()
