import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import synth
import vision_slam_frontend_b200 as vsf
from oracle import native, restate
ratio = restate.NN_MATCH_RATIO
with vsf.Context(device=0, max_features=4096, desc_bytes=32, window=3) as ctx:
    for eng in (2, 3, 1):
        ctx.set_engine(eng, 0)
        for (nq, nt, seed) in [(2300, 2100, 2), (257, 4001, 3), (1000, 33, 4), (3000, 3000, 5)]:
            Q, T = synth.descriptor_pair(nq, nt, seed=seed)
            idx, dist = ctx.knn2(Q, T)
            ei, ed = native.knn2_hamming(Q, T)
            assert (idx == ei).all() and (dist == ed).all()
            assert (ctx.get_matches(Q, T, ratio) == native.get_matches(Q, T, ratio)).all()
        for split in (1, 3, 7):
            ctx.set_tuning(-1, split, 0, -1)
            Q, T = synth.descriptor_pair(1500, 1700, seed=split)
            idx, dist = ctx.knn2(Q, T)
            ei, ed = native.knn2_hamming(Q, T)
            assert (idx == ei).all() and (dist == ed).all()
        ctx.set_tuning()
    ctx.set_engine(2, 0)
    frames = [synth.synth_pose(1300 - 53 * (p % 4), p, 130, 17) for p in range(12)]
    for p, D in enumerate(frames):
        if ctx.window_in_flight() == vsf.PIPELINE_DEPTH:
            ctx.window_collect()
        ctx.window_submit(100 + p, D, ratio, 0.3, p % 2, False)
    while ctx.window_in_flight():
        ctx.window_collect()
print("sanitizer target ok")
