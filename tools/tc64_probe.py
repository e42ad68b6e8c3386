"""Bring-up check of the 64-byte tensor-core engine against the oracle (run under gpurun, wrapped in timeout)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
import vision_slam_frontend_b200 as vsf
from oracle import native, restate
ratio = restate.NN_MATCH_RATIO
bad = 0
for width in (64, 61):
    with vsf.Context(device=0, max_features=6000, desc_bytes=width, window=3) as ctx:
        ctx.set_engine(2, 0)
        for (nq, nt, seed) in [(128, 128, 1), (700, 650, 2), (129, 300, 3), (2300, 2100, 4), (1000, 33, 5), (77, 1, 6), (5000, 5000, 7), (300, 4097, 8)]:
            Q, T = synth.descriptor_pair(nq, nt, width=width, seed=seed)
            idx, dist = ctx.knn2(Q, T)
            ei, ed = native.knn2_hamming(Q, T)
            ok = bool((idx == ei).all() and (dist == ed).all())
            gm = ctx.get_matches(Q, T, ratio); em = native.get_matches(Q, T, ratio)
            okm = bool(len(gm) == len(em) and (gm == em).all())
            print(width, nq, nt, "engine", ctx.last_engine, "knn", ok, "matches", okm, len(em), flush=True)
            if not (ok and okm):
                bad += 1
                w = np.nonzero((idx != ei).any(1) | (dist != ed).any(1))[0][:3]
                for i in w: print("   q", i, "exp", ei[i], ed[i], "got", idx[i], dist[i])
        for split in (1, 2, 5):
            ctx.set_tuning(-1, split, 0, -1)
            Q, T = synth.descriptor_pair(1500, 1700, width=width, seed=split)
            idx, dist = ctx.knn2(Q, T)
            ei, ed = native.knn2_hamming(Q, T)
            ok = bool((idx == ei).all() and (dist == ed).all())
            print(width, "split", split, ok, flush=True)
            bad += 0 if ok else 1
        ctx.set_tuning()
        Q, T = synth.tie_pair(900, 1100) if hasattr(synth, "tie_pair") else (None, None)
print("BAD", bad)
