"""Bring-up / timing probe for the tensor-core kNN engine (run on the GPU box under gpurun).

`python tools/tc_probe.py`          drives every (engine, flags) case in its own subprocess
                                    with a timeout, so a hung kernel cannot take the call down.
`python tools/tc_probe.py case E F` runs one case: parity of engine E (flags F) against the
                                    POPC engine on the same device, then timings on the C4 shape.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_case(engine: int, flags: int, timing: bool):
    import ctypes as C

    import numpy as np
    import torch

    import synth
    import vision_slam_frontend_b200 as vsf

    out = {"engine": engine, "flags": flags, "parity": []}
    ctx = vsf.Context(device=0, max_features=20000, desc_bytes=32, window=10)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    shapes = [(700, 650, 0), (256, 256, 1), (2000, 2000, 2), (5000, 5000, 3), (300, 33, 4), (77, 1, 5),
              (1000, 4097, 6)]
    for (nq, nt, seed) in shapes:
        Q, T = synth.descriptor_pair(nq, nt, seed=seed)
        ctx.set_engine(1, 0)
        ri, rd = ctx.knn2(Q, T)
        rm = ctx.get_matches(Q, T, float(np.float32(0.6)))
        ctx.set_engine(engine, flags)
        gi, gd = ctx.knn2(Q, T)
        gm = ctx.get_matches(Q, T, float(np.float32(0.6)))
        used = ctx.last_engine
        bad_i = int((ri != gi).any(axis=1).sum())
        bad_d = int((rd != gd).any(axis=1).sum())
        ok_m = bool(len(rm) == len(gm) and (rm == gm).all())
        rec = dict(nq=nq, nt=nt, used=used, bad_idx_rows=bad_i, bad_dist_rows=bad_d, matches_equal=ok_m,
                   n_matches=int(len(rm)))
        if bad_d:
            w = np.nonzero((rd != gd).any(axis=1))[0][:4]
            rec["sample"] = [dict(q=int(i), ref=[ri[i].tolist(), rd[i].tolist()], got=[gi[i].tolist(), gd[i].tolist()])
                             for i in w]
        out["parity"].append(rec)
        print(rec, flush=True)
    # ties
    Q, T = synth.tie_pair(900, 1100)
    ctx.set_engine(1, 0)
    ri, rd = ctx.knn2(Q, T)
    ctx.set_engine(engine, flags)
    gi, gd = ctx.knn2(Q, T)
    out["ties_equal"] = bool((ri == gi).all() and (rd == gd).all())
    print("ties_equal", out["ties_equal"], flush=True)

    if timing:
        RATIO = float(np.float32(0.6))

        def time_window(n, W, poses, reps=3):
            seq = torch.empty((poses + W, n, 32), dtype=torch.uint8, device="cuda")
            ctx.synth_sequence_device(seq.data_ptr(), n, 0, poses + W, max(1, n // 10), 7)
            base, fb, L = seq.data_ptr(), n * 32, ctx._L

            def step(t):
                qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
                nn = (C.c_int * W)(*([n] * W))
                rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO)
                assert rc == 0, L.vsf_last_error(ctx._h)
            for t in range(5):
                step(t)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for t in range(poses):
                    step(t)
                e1.record(stream)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / poses)
            counts = np.zeros(W, np.int32)
            ctx._check(L.vsf_fetch_window(ctx._h, W, counts.ctypes.data, None, 0))
            return best, counts.tolist()

        out["timing"] = []
        for (n, W, poses) in [(5000, 10, 64), (20000, 10, 4), (2000, 1, 100)]:
            for eng in (1, engine):
                for split in ((0,) if eng == 1 else (0, 1, 2, 3, 4, 5, 8)):
                    ctx.set_engine(eng, flags if eng != 1 else 0)
                    ctx.set_tuning(-1, split, 0, -1)
                    ms, counts = time_window(n, W, poses)
                    r = dict(n=n, W=W, engine=eng, split=split, ms=ms, gcmp_s=W * n * n / (ms * 1e-3) / 1e9,
                             counts=counts[:3])
                    out["timing"].append(r)
                    print(r, flush=True)
    return out


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if len(sys.argv) >= 4 and sys.argv[1] == "case":
        e, f = int(sys.argv[2]), int(sys.argv[3])
        timing = len(sys.argv) >= 5 and sys.argv[4] == "time"
        res = run_case(e, f, timing)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tc_probe_e{e}_f{f}.json"), "w"), indent=1)
        return
    summary = {}
    good = []
    for e in (2, 3):
        for f in (0, 1):
            try:
                p = subprocess.run([sys.executable, __file__, "case", str(e), str(f)], timeout=240,
                                   capture_output=True, text=True)
                tail = (p.stdout + p.stderr)[-3000:]
                summary[f"e{e}_f{f}"] = dict(rc=p.returncode, tail=tail)
                path = os.path.join(ROOT, "gpurun_out", f"tc_probe_e{e}_f{f}.json")
                if p.returncode == 0 and os.path.exists(path):
                    r = json.load(open(path))
                    if all(x["bad_dist_rows"] == 0 and x["bad_idx_rows"] == 0 and x["matches_equal"] for x in r["parity"]) \
                            and r["ties_equal"]:
                        good.append((e, f))
            except subprocess.TimeoutExpired as ex:
                summary[f"e{e}_f{f}"] = dict(rc="timeout", tail=str(ex.stdout)[-1500:] if ex.stdout else "")
            print(f"case e{e} f{f}:", summary[f"e{e}_f{f}"]["rc"], flush=True)
            print(summary[f"e{e}_f{f}"]["tail"][-1200:], flush=True)
    summary["good"] = good
    print("GOOD CASES:", good, flush=True)
    for (e, f) in good:
        try:
            p = subprocess.run([sys.executable, __file__, "case", str(e), str(f), "time"], timeout=600,
                               capture_output=True, text=True)
            summary[f"time_e{e}_f{f}"] = dict(rc=p.returncode, tail=(p.stdout + p.stderr)[-6000:])
            print(summary[f"time_e{e}_f{f}"]["tail"], flush=True)
        except subprocess.TimeoutExpired:
            summary[f"time_e{e}_f{f}"] = dict(rc="timeout")
    json.dump(summary, open(os.path.join(ROOT, "gpurun_out", "tc_probe_summary.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
