"""Timing probe for the tensor-core engine's main kernel (run under gpurun): per-kernel CUDA-event
times (vsf_set_profile) of the window launch for a list of (features, window) shapes and
bring-up flags (2 = no bucket reduction, 4 = no TMEM loads; they only act in a library built with
`VSF_TC_BRINGUP=1 python -m vision_slam_frontend_b200.build --force`), to separate the fixed cost
of a launch from the per-tile cost.

    python tools/tc_scaling.py [out.json]
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

    RATIO = float(np.float32(0.6))
    out = []
    ctx = vsf.Context(device=0, max_features=20480, desc_bytes=32, window=10)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = ctx._L
    shapes = [(5000, 10), (5120, 10), (4864, 10), (2560, 10), (10240, 10), (20480, 10), (5000, 1), (1024, 10)]
    for (n, W) in shapes:
        poses = 24
        seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
        base, fb = seq.data_ptr(), n * 32

        def step(t):
            qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
            nn = (C.c_int * W)(*([n] * W))
            rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
            assert rc == 0, L.vsf_last_error(ctx._h)

        for flags in (0, 2, 4):
            ctx.set_engine(2, flags)
            ctx.set_profile(True)
            for t in range(4):
                step(t)
            acc = np.zeros(4)
            for t in range(poses):
                step(t)
                acc += np.array(ctx.last_kernel_times())
            acc /= poses
            ctx.set_profile(False)
            # chained (programmatic dependent launch) time of the whole sequence
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(4):
                for t in range(poses):
                    step(t)
            e1.record(stream)
            torch.cuda.synchronize()
            chain = e0.elapsed_time(e1) / (4 * poses)
            tiles = ((n + 255) // 256) ** 2 * W
            rec = dict(n=n, W=W, flags=flags, expand_ms=acc[0], main_ms=acc[1], refine_ms=acc[2], compact_ms=acc[3],
                       chain_ms=chain, tile_slots=tiles, tile_slots_per_sm=tiles / 148.0,
                       us_per_tile_slot=acc[1] * 1e3 / (tiles / 148.0), gcmp_s=W * n * n / (chain * 1e-3) / 1e9)
            out.append(rec)
            print(json.dumps(rec), flush=True)
        del seq
    ctx.set_engine(0, 0)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tc_scaling.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
