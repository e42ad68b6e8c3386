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
# the fused ObserveImage path: stereo match + residual filter + compaction, window, triangulation
P1, P2 = synth.kitti_projections()
F = synth.kitti_fundamental()
for eng in (0, 2):
    with vsf.Context(device=0, max_features=4096, desc_bytes=32, window=3) as ctx:
        ctx.set_engine(eng, 0)
        for p, (kl, dl, kr, dr) in enumerate(synth.stereo_sequence(4, 1800 if eng else 500, seed=6)):
            ctx.observe_features(p, kl, dl, kr, dr, F, P1, P2, ratio)
# blocking window calls with the device sort and the host sort; 64-byte descriptors (POPC engine)
with vsf.Context(device=0, max_features=2048, desc_bytes=32, window=3) as ctx:
    for p in range(5):
        D = synth.synth_pose(1500 - 31 * p, p, 150, 3)
        for mode in (0, 1):
            ctx.window_feature_matches(D, ratio, 0.3, mode)
        ctx.window_match(D, ratio)
        ctx.window_commit(p, len(D))
with vsf.Context(device=0, max_features=1024, desc_bytes=61, window=2) as ctx:
    Q, T = synth.descriptor_pair(700, 650, width=61, seed=8)
    idx, dist = ctx.knn2(Q, T)
    ei, ed = native.knn2_hamming(Q, T)
    assert (idx == ei).all() and (dist == ed).all()
x1 = np.random.default_rng(0).random((300, 2), dtype=np.float32) * 300
with vsf.Context(device=0, max_features=1024, desc_bytes=32, window=2) as ctx:
    ctx.triangulate(P1, P2, x1, x1 + np.float32([5, 0]))
    ctx.undistort_points(np.float32([[700, 0, 640], [0, 705, 512], [0, 0, 1]]), np.float32([-0.3, 0.1, 1e-3, -2e-3, 0]), x1)
# 61-byte descriptors on the tensor cores (knn2_tc64_kernel.cu)
with vsf.Context(device=0, max_features=2600, desc_bytes=61, window=2) as ctx:
    ctx.set_engine(2, 0)
    for (nq, nt, seed) in [(2300, 2100, 2), (129, 300, 3), (1000, 33, 4)]:
        Q, T = synth.descriptor_pair(nq, nt, width=61, seed=seed)
        idx, dist = ctx.knn2(Q, T)
        ei, ed = native.knn2_hamming(Q, T)
        assert (idx == ei).all() and (dist == ed).all()
# round 2: the CTA-pair kernel (enough work for every pair: 3600 x 3600 at 61 bytes), both refine variants
with vsf.Context(device=0, max_features=3600, desc_bytes=61, window=2) as ctx:
    ctx.set_engine(2, 0)
    Q, T = synth.descriptor_pair(3600, 3500, width=61, seed=12)
    idx, dist = ctx.knn2(Q, T)
    ei, ed = native.knn2_hamming(Q, T)
    assert (idx == ei).all() and (dist == ed).all()
    assert (ctx.get_matches(Q, T, ratio) == native.get_matches(Q, T, ratio)).all()
# round 2: the exact device sort (all three scratch paths: short list, shared-memory scratch, heapsort fallback),
# the side-stream sort of the pipelined path, the pipelined whole-frame path
from vision_slam_frontend_b200 import capi
with vsf.Context(device=0, max_features=4096, desc_bytes=32, window=3) as ctx:
    rng = np.random.default_rng(1)
    for n in (10, 300, 4000):
        m = np.zeros(n, vsf.DMATCH_DTYPE)
        m["queryIdx"] = np.arange(n); m["trainIdx"] = np.arange(n); m["distance"] = rng.integers(0, 30, n).astype(np.float32)
        ctx.debug_sort_device(m, 0.3, True)
    ctx.set_option(capi.OPT_DEBUG_SORT_DEPTH, 1)
    ctx.debug_sort_device(m, 0.3, True)
    ctx.set_option(capi.OPT_DEBUG_SORT_DEPTH, -1)
    frames = [synth.synth_pose(1300 - 53 * (p % 4), p, 130, 17) for p in range(10)]
    for p, D in enumerate(frames):
        if ctx.window_in_flight() == vsf.PIPELINE_DEPTH:
            ctx.window_collect()
        ctx.window_submit(100 + p, D, ratio, 0.3, 2, False)
    while ctx.window_in_flight():
        ctx.window_collect()
K = synth.KITTI_K.astype(np.float32)
with vsf.Context(device=0, max_features=2048, desc_bytes=32, window=3) as ctx:
    for p, (kl, dl, kr, dr) in enumerate(synth.stereo_sequence(6, 900, seed=7)):
        ctx.observe_submit(p, kl, dl, kr, dr, F, P1, P2, K, np.zeros(5, np.float32), ratio)
        if ctx.observe_in_flight() >= 3:
            ctx.observe_collect()
    while ctx.observe_in_flight():
        ctx.observe_collect()
with vsf.Context(device=0, max_features=20000, desc_bytes=32, window=2) as ctx:   # exact sort with global scratch
    m = np.zeros(18000, vsf.DMATCH_DTYPE)
    m["queryIdx"] = np.arange(18000); m["distance"] = np.random.default_rng(2).integers(0, 40, 18000).astype(np.float32)
    ctx.debug_sort_device(m, 0.3, True)
# round 2 (second half): groups of poses - early-start distance kernels, finish kernels side by side (both
# widths), the host-buffer sequence call with grouped launches, every sort mode
import torch
for width, n, W in ((32, 1300, 3), (61, 1500, 2)):
    with vsf.Context(device=0, max_features=2048, desc_bytes=width, window=W) as ctx:
        ctx.set_engine(2, 0)
        rb = ctx.row_bytes
        poses = W + 9
        buf = torch.empty((poses, n, rb), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(buf.data_ptr(), n, 0, poses, 97, 5)
        for group, count in ((4, 9), (1, 3), (3, 4)):
            ctx.set_option(capi.OPT_POSE_GROUP, group)
            ctx.window_match_block_device(buf.data_ptr(), n, poses, 1, count, ratio)
            ctx.fetch_window(W)
        ctx.set_option(capi.OPT_POSE_GROUP, 4)
        host = torch.empty((poses, n, rb), dtype=torch.uint8).pin_memory()
        host.copy_(buf)
        torch.cuda.synchronize()
        hp = host.numpy()
        for p in range(W):
            ctx.window_push(p, hp[p][:, :width])
        out = np.zeros((4, W, n), capi.FEATURE_MATCH_DTYPE)
        cnt = np.zeros((4, W), np.int32)
        for mode in (1, 2, 0):
            ctx.window_run_sequence(hp, W, 13, ratio, 0.3, mode, 12, out, cnt)
print("sanitizer target ok")
