"""Where the automatic engine choice should switch from the POPC kernel to the tensor-core
sequence: back-to-back launch time of both engines on small window shapes (run under gpurun)."""
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

    RATIO = float(np.float32(0.6))
    ctx = vsf.Context(device=0, max_features=8192, desc_bytes=32, window=10)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    L = ctx._L
    out = []
    for (n, W) in [(1000, 1), (2000, 1), (3000, 1), (4000, 1), (5000, 1), (8000, 1), (1000, 10), (2000, 10), (1000, 4), (2000, 4)]:
        poses = 32
        seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
        base, fb = seq.data_ptr(), n * 32

        def step(t):
            qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
            nn = (C.c_int * W)(*([n] * W))
            rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
            assert rc == 0

        rec = dict(n=n, W=W, cmp=W * n * n)
        for eng, name in ((1, "popc_us"), (2, "tensor_us")):
            ctx.set_engine(eng, 0)
            for t in range(8):
                step(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(8):
                for t in range(poses):
                    step(t)
            e1.record(stream)
            torch.cuda.synchronize()
            rec[name] = 1e3 * e0.elapsed_time(e1) / (8 * poses)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    ctx.set_engine(0, 0)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "engine_crossover.json")
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
