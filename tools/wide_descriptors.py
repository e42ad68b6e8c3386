"""N4 (descriptor widths 61 / 64 bytes: AKAZE is the reference's default extractor,
src/slam_frontend.cc:553): throughput of both engines on 64-byte rows for the C4 and C2
shapes, next to OpenCV's BFMatcher on the same arrays (run under gpurun)."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import numpy as np
    import torch

    import vision_slam_frontend_b200 as vsf
    from oracle import cv2_ref, native

    RATIO = float(np.float32(0.6))
    out = []
    for (n, W, width) in [(5000, 10, 64), (5000, 10, 61), (2000, 1, 61)]:
        ctx = vsf.Context(device=0, max_features=n, desc_bytes=width, window=W)
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        ctx.set_stream(stream.cuda_stream)
        L = ctx._L
        poses = 12
        g = torch.Generator(device="cuda").manual_seed(3)
        seq = torch.randint(0, 256, (poses + W, n, 64), dtype=torch.uint8, device="cuda", generator=g)
        if width < 64:
            seq[:, :, width:] = 0            # device rows are zero-padded to 64 bytes
        # plant matches: 40 % of every frame's rows repeat a row of the previous frame with a few flipped bits
        k = int(0.4 * n)
        for p in range(1, poses + W):
            src = torch.randperm(n, device="cuda", generator=g)[:k]
            flips = torch.zeros((k, 64), dtype=torch.uint8, device="cuda")
            flips[:, : width // 8] = torch.randint(0, 2, (k, width // 8), dtype=torch.uint8, device="cuda", generator=g)
            seq[p, :k] = seq[p - 1, src] ^ flips
        torch.cuda.synchronize()
        base, fb = seq.data_ptr(), n * 64

        def step(t):
            qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
            nn = (C.c_int * W)(*([n] * W))
            rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
            assert rc == 0, L.vsf_last_error(ctx._h)

        times = {}
        for eng, name in ((1, "popc"), (2, "tensor")):
            ctx.set_engine(eng, 0)
            for t in range(4):
                step(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(4):
                for t in range(poses):
                    step(t)
            e1.record(stream)
            torch.cuda.synchronize()
            times[name] = e0.elapsed_time(e1) / (4 * poses)
        ms = times["tensor"]
        ctx.set_profile(True)                      # per-kernel event times of the tensor sequence
        kt = np.zeros(4)
        for t in range(poses):
            step(t)
            kt += np.array(ctx.last_kernel_times())
        ctx.set_profile(False)
        kt /= poses
        got = ctx.fetch_window(W)
        # parity of the last pose's newest pair against the oracle, and the CPU time of that pair
        host = seq[poses - 1 + W - 1: poses + W].cpu().numpy()[:, :, :width].copy()
        t_last = poses - 1
        exp = native.get_matches(host[0], host[1], RATIO)
        same = bool(len(got[W - 1]) == len(exp) and (got[W - 1] == exp).all())
        cpu_s = None
        if cv2_ref.available():
            cv2_ref.set_threads(os.cpu_count() or 1)
            cv2_ref.knn_match_raw(host[0], host[1])
            t0 = time.perf_counter()
            cv2_ref.knn_match_raw(host[0], host[1])
            cpu_s = time.perf_counter() - t0
        rec = dict(features=n, window=W, descriptor_bytes=width, engine=ctx.last_engine, us_per_pose=ms * 1e3,
                   gpu_cmp_per_s=W * n * n / (ms * 1e-3), popc_us_per_pose=times["popc"] * 1e3,
                   popc_cmp_per_s=W * n * n / (times["popc"] * 1e-3),
                   tensor_kernel_us=dict(expand=kt[0] * 1e3, distance=kt[1] * 1e3, refine=kt[2] * 1e3, compact=kt[3] * 1e3), survivors_newest_pair=int(len(exp)), bit_exact_vs_oracle=same,
                   cpu_pair_s=cpu_s, cpu_cmp_per_s=(n * n / cpu_s) if cpu_s else None, cpu_threads=os.cpu_count())
        out.append(rec)
        print(json.dumps(rec), flush=True)
        ctx.close()
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "wide_descriptors.json")
    json.dump(out, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
