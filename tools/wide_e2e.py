"""End-to-end rate of the pipelined frame stream (vsf_window_submit / vsf_window_collect, host
buffers, exact sort order) on 61-byte (AKAZE) descriptors, C4 shape (run under gpurun)."""
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

    n, W, width, poses = 5000, 10, 61, 160
    RATIO = float(np.float32(0.6))
    g = torch.Generator(device="cuda").manual_seed(3)
    seq = torch.randint(0, 256, (poses + W, n, width), dtype=torch.uint8, device="cuda", generator=g)
    k = int(0.4 * n)
    for p in range(1, poses + W):        # 40 % of every frame repeats rows of the previous one, a few bits flipped
        src = torch.randperm(n, device="cuda", generator=g)[:k]
        flips = torch.zeros((k, width), dtype=torch.uint8, device="cuda")
        flips[:, :7] = torch.randint(0, 2, (k, 7), dtype=torch.uint8, device="cuda", generator=g)
        seq[p, :k] = seq[p - 1, src] ^ flips
    host = seq.cpu().pin_memory()
    hp = host.numpy()
    out = {}
    for engine, name in ((2, "tensor"), (1, "popc")):
        ctx = vsf.Context(device=0, max_features=n, desc_bytes=width, window=W)
        ctx.set_engine(engine, 0)
        L = ctx._L
        for p in range(W):
            ctx.window_push(p, hp[p])
        fids = np.zeros(W, np.uint64); counts = np.zeros(W, np.int32)
        fm = np.zeros((W, n), dtype=vsf.FEATURE_MATCH_DTYPE); nf = C.c_int(0); fid = C.c_uint64(0)
        lag = vsf.PIPELINE_DEPTH - 2
        kept = 0

        def collect():
            nonlocal kept
            rc = L.vsf_window_collect(ctx._h, C.byref(fid), fids.ctypes.data, counts.ctypes.data, fm.ctypes.data, n, C.byref(nf))
            assert rc == 0, L.vsf_last_error(ctx._h)
            kept += int(counts[: nf.value].sum())

        def run(first, count):
            for t in range(first, first + count):
                D = hp[W + t]
                rc = L.vsf_window_submit(ctx._h, W + t, D.ctypes.data, n, width, RATIO, 0.3, 1, 0)
                assert rc == 0, L.vsf_last_error(ctx._h)
                if L.vsf_window_in_flight(ctx._h) > lag:
                    collect()
            while L.vsf_window_in_flight(ctx._h):
                collect()

        run(0, 30)
        torch.cuda.synchronize()
        kept = 0
        t0 = time.perf_counter()
        run(30, poses - 30)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        out[name] = dict(us_per_pose=1e6 * el / (poses - 30), cmp_per_s=(poses - 30) * W * n * n / el,
                         feature_matches_per_pose=kept / (poses - 30), engine=ctx.last_engine)
        print(name, json.dumps(out[name]), flush=True)
        ctx.close()
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "wide_e2e.json")
    json.dump(dict(features=n, window=W, descriptor_bytes=width, sort_mode=1, results=out), open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
