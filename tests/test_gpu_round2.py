"""Parity cases added in round 2: the shapes and switches the first round left unchecked.

* a5 residual: both float32 summation orders Eigen may use, each bit-identical to the oracle;
* the adaptive threshold through a frame without stereo matches (reference NaN / opt-in hold);
* vsf_observe_features when frame sizes shrink between poses (capacity contract);
* C5's real shape: 20000 features against a 32-frame window, every list checked;
* the C++ Frontend mirror at 61-byte (AKAZE, the reference default) descriptors;
* a sequence sharded into pose ranges with a (window + 1)-frame halo reproduces the unsharded
  output bit for bit (two contexts: on two GPUs when the box has them, else on one).
"""
import ctypes as C

import numpy as np
import pytest

import synth
from oracle import native, restate

pytestmark = pytest.mark.gpu

RATIO = restate.NN_MATCH_RATIO
BP = restate.BEST_PERCENT


def same_f32(a, b) -> bool:
    """Bit-identical float32 values; any NaN equals any NaN (the payload of 0/0 is the FPU's)."""
    a, b = np.float32(a), np.float32(b)
    return bool((np.isnan(a) and np.isnan(b)) or a.view(np.uint32) == b.view(np.uint32))


def new_ctx(**kw):
    import vision_slam_frontend_b200 as vsf
    args = dict(device=0, max_features=4096, desc_bytes=32, window=10)
    args.update(kw)
    return vsf.Context(**args)


# ------------------------------------------------------------------ a5: residual order switch

@pytest.mark.parametrize("order", [0, 1])
def test_stereo_residual_order_switch(order):
    from vision_slam_frontend_b200 import capi
    # a rectified rig's F has six zero entries, which makes every partial sum exact; perturb
    # all nine so that the two summation orders have something to disagree about
    F = (synth.kitti_fundamental() + np.random.default_rng(2).normal(0, 2e-4, (3, 3))).astype(np.float32)
    frames = synth.stereo_sequence(3, 2000, seed=11)
    thresh = restate.STEREO_AMBIG_INIT
    differs = 0
    with new_ctx() as ctx:
        ctx.set_option(capi.OPT_RESIDUAL_ORDER, order)
        assert ctx.get_option(capi.OPT_RESIDUAL_ORDER) == order
        for kl, dl, kr, dr in frames:
            got = ctx.stereo_filter(kl, dl, kr, dr, F, RATIO)
            sm = native.get_matches(dl, dr, RATIO)
            L, R, thresh_next, c, keep = restate.remove_ambig_stereo(
                restate.Frame(kl, dl, 0), restate.Frame(kr, dr, 0), sm, F, thresh, residual_order=order)
            np.testing.assert_array_equal(got["residuals"].view(np.uint32), c.view(np.uint32))
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][keep])
            assert ctx.get_stereo_threshold().view(np.uint32) == np.float32(thresh_next).view(np.uint32)
            # what is actually known about the other order: at most one ulp of the accumulated
            # sum away (relative to the residual that can be more after cancellation), and it
            # only matters for a pair sitting exactly on the threshold
            qi, ti = sm["queryIdx"], sm["trainIdx"]
            xl = np.stack([kl["x"][qi], kl["y"][qi]], 1)
            xr = np.stack([kr["x"][ti], kr["y"][ti]], 1)
            other = restate.epipolar_residual(xl, xr, F, 1 - order)
            differs += int((other.view(np.uint32) != c.view(np.uint32)).sum())
            scale = np.abs(xl).max() * np.abs(F).max() * np.abs(xr).max() * 4
            assert np.abs(other.astype(np.float64) - c.astype(np.float64)).max() <= 4 * np.spacing(np.float32(scale))
            thresh = thresh_next
    assert differs > 0, "the two orders never differed on this data: the switch is not exercised"


@pytest.mark.parametrize("hold", [0, 1])
def test_threshold_through_an_empty_frame(hold):
    """Reference arithmetic (src/slam_frontend.cc:392-394): 0 matches -> 0/0 + 2 = NaN, and every
    later frame loses all its stereo pairs.  VSF_OPT_HOLD_THRESHOLD_ON_EMPTY = 1 keeps the old
    threshold instead (opt-in deviation)."""
    from vision_slam_frontend_b200 import capi
    F = synth.kitti_fundamental()
    P1, P2 = synth.kitti_projections()
    frames = synth.stereo_sequence(3, 800, seed=13)
    rng = np.random.default_rng(0)
    kl = synth.make_keypoints(rng.uniform(0, 300, (60, 2)))
    empty = (kl, rng.integers(0, 256, (60, 32), dtype=np.uint8), kl, rng.integers(0, 256, (60, 32), dtype=np.uint8))
    seq = [frames[0], empty, frames[1], frames[2]]
    fo = restate.FrontendOracle(P1, P2, F, frame_life=3, order="stable", hold_on_empty=bool(hold))
    with new_ctx(window=3) as ctx:
        ctx.set_option(capi.OPT_HOLD_THRESHOLD_ON_EMPTY, hold)
        for p, (a, b, c, d) in enumerate(seq):
            r = fo.observe_features(a, b, c, d)
            got = ctx.observe_features(p, a, b, c, d, F, P1, P2, RATIO)
            np.testing.assert_array_equal(got["kept_left"], r.stereo_matches["queryIdx"][r.stereo_keep])
            assert same_f32(got["stereo_threshold_next"], fo.stereo_ambig_constraint)
            if p == 1:
                assert np.isnan(got["stereo_threshold_next"]) == (hold == 0)
            if p == 2:     # the frame filtered with that threshold: emptied, or untouched
                assert (len(got["kept_left"]) > 300) if hold else (len(got["kept_left"]) == 0)
        # mean + 2 depends only on the previous frame's raw matches, so the NaN lasts one frame
        assert len(got["kept_left"]) > 300


# ------------------------------------------------- capacity contract of vsf_observe_features

def _observe_raw(ctx, frame_id, kl, dl, kr, dr, F, P1, P2, cap):
    """vsf_observe_features with a caller-chosen capacity (the Python wrapper always passes
    max_features)."""
    from vision_slam_frontend_b200 import capi
    kl = np.ascontiguousarray(kl, capi.KEYPOINT_DTYPE)
    kr = np.ascontiguousarray(kr, capi.KEYPOINT_DTYPE)
    W = ctx.window
    kept_l, kept_r = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    fids, wc = np.zeros(W, np.uint64), np.zeros(W, np.int32)
    guard = 64
    wm = np.zeros(W * cap + guard, capi.DMATCH_DTYPE)
    wm["queryIdx"][W * cap:] = -777                     # canary behind the caller's buffer
    tm = np.zeros(cap, capi.DMATCH_DTYPE)
    X4 = np.zeros((cap, 4), np.float32)
    o = capi._ObserveOut(kept_l.ctypes.data, kept_r.ctypes.data, 0, 0.0, 0, fids.ctypes.data, wc.ctypes.data,
                         wm.ctypes.data, 0, tm.ctypes.data, X4.ctypes.data, cap)
    Ff = np.ascontiguousarray(F, np.float32).reshape(9)
    P1f = np.ascontiguousarray(P1, np.float32).reshape(12)
    P2f = np.ascontiguousarray(P2, np.float32).reshape(12)
    rc = ctx._L.vsf_observe_features(ctx._h, frame_id, kl.ctypes.data, dl.ctypes.data, len(kl), dl.strides[0],
                                     kr.ctypes.data, dr.ctypes.data, len(kr), dr.strides[0], Ff.ctypes.data,
                                     P1f.ctypes.data, P2f.ctypes.data, float(RATIO), C.byref(o))
    assert (wm["queryIdx"][W * cap:] == -777).all(), "wrote behind the caller's window_matches buffer"
    lists = [wm[j * cap: j * cap + wc[j]].copy() for j in range(o.n_frames)] if rc == 0 else None
    return rc, lists, kept_l[:o.n_kept].copy()


def test_observe_features_when_frames_shrink():
    """A past frame with ~1700 kept rows followed by a 300-feature frame: the window list of that
    pair can hold more entries than the current frame has features.  A capacity of n_left is
    refused (VSF_ERR_CAPACITY, state untouched); with enough room the lists are the oracle's."""
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    big = synth.stereo_sequence(2, 2000, seed=31)
    # the small frame re-observes 300 of the second big frame's landmarks (same descriptors, few flips)
    rng = np.random.default_rng(5)
    kl2, dl2, kr2, dr2 = big[1]
    rows = rng.permutation(len(kl2))[:300]
    small = (kl2[rows], synth.flip_bits(dl2[rows], 4, rng), kr2[rows], synth.flip_bits(dr2[rows], 4, rng))
    fo = restate.FrontendOracle(P1, P2, F, frame_life=3, order="stable")
    with new_ctx(window=3) as ctx:
        for p in range(2):
            fo.observe_features(*big[p])
            rc, _, _ = _observe_raw(ctx, p, *big[p], F, P1, P2, cap=2000)
            assert rc == 0
        thr_before = ctx.get_stereo_threshold()
        rc, _, _ = _observe_raw(ctx, 2, *small, F, P1, P2, cap=300)
        assert rc == 2                                   # VSF_ERR_CAPACITY
        assert ctx.window_size() == 2                    # nothing was committed ...
        assert ctx.get_stereo_threshold().view(np.uint32) == thr_before.view(np.uint32)   # ... or advanced
        past = list(fo.frame_list)
        r = fo.observe_features(*small)
        rc, lists, kept = _observe_raw(ctx, 2, *small, F, P1, P2, cap=2000)
        assert rc == 0 and ctx.window_size() == 3
        np.testing.assert_array_equal(kept, r.stereo_matches["queryIdx"][r.stereo_keep])
        for m, pf in zip(lists, past):
            np.testing.assert_array_equal(m, native.get_matches(pf.descriptors, r.left.descriptors, RATIO))
        assert ctx.get_stereo_threshold().view(np.uint32) == np.float32(fo.stereo_ambig_constraint).view(np.uint32)


def test_descriptor_width_is_checked():
    with new_ctx(desc_bytes=61) as ctx:
        Q, T = synth.descriptor_pair(100, 100, width=64, seed=1)
        with pytest.raises(ValueError):
            ctx.knn2(Q, T)


# ----------------------------------------------------------------- C5 at its real shape

def test_window_c5_real_shape():
    """BASELINE config 5's per-pose shape: 20000 features against a 32-frame window, 33 problems
    in one launch sequence (kMaxProblems = 40), through the blocking host call with the
    reference's sort order; every one of the 32 lists is compared with the oracle."""
    n, W = 20000, 32
    stride = n // W
    frames = [synth.synth_pose(n, p, stride, 23) for p in range(W + 2)]
    with new_ctx(max_features=n, window=W) as ctx:
        for p in range(W):
            ctx.window_push(p, frames[p])
        got = ctx.window_feature_matches(frames[W], RATIO, float(BP), sort_mode=1)
        assert ctx.last_engine == 2
        assert len(got) == W
        total = 0
        for j, (fid, pairs) in enumerate(got):
            assert fid == j
            m = native.get_matches(frames[j], frames[W], RATIO)
            keep = restate.num_good_matches(len(m), BP)
            exp = m[restate.sort_order_stdsort(m)][:keep]
            assert len(pairs) == keep
            np.testing.assert_array_equal(pairs[:, 0], exp["queryIdx"].astype(np.uint64))
            np.testing.assert_array_equal(pairs[:, 1], exp["trainIdx"].astype(np.uint64))
            total += keep
        assert total > 30000
        # eviction at W = 32 (src/slam_frontend.cc:467-470), then the query-ordered lists of the
        # next pose for the oldest and the newest resident frame
        ctx.window_commit(W, n)
        assert ctx.window_size() == W
        got = ctx.window_match(frames[W + 1], RATIO)
        assert [fid for fid, _ in got] == list(range(1, W + 1))
        for j in (0, W - 1):
            np.testing.assert_array_equal(got[j][1], native.get_matches(frames[j + 1], frames[W + 1], RATIO))


# ------------------------------------------- the C++ Frontend at the reference's default width

@pytest.mark.parametrize("exact", [True, False])
def test_frontend_sequence_61_byte_descriptors(exact):
    from vision_slam_frontend_b200.frontend import Frontend
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    W = 3
    frames = synth.stereo_sequence(5, 1500, seed=41, width=61)
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order="stdsort" if exact else "stable")
    with Frontend(max_features=2048, desc_bytes=61, frame_life=W, P_left=P1, P_right=P2, fundamental=F,
                  exact_std_sort=exact) as fe:
        for p, (kl, dl, kr, dr) in enumerate(frames):
            fe.observe_odometry([0.5 * (p + 1), 0, 0], [1, 0, 0, 0], 10.0 + p)
            assert fe.observe_features(kl, dl, kr, dr, 10.0 + p)
            fo.observe_features(kl, dl, kr, dr)
            assert fe.stereo_threshold.view(np.uint32) == np.float32(fo.stereo_ambig_constraint).view(np.uint32)
        got, exp = fe.vision_factors(), fo.vision_factors
        assert len(got) == len(exp)
        for (a, b, pairs), e in zip(got, exp):
            assert (a, b) == (e.pose_idx_initial, e.pose_idx_current)
            np.testing.assert_array_equal(pairs, e.feature_matches)
        assert sum(len(p) for _, _, p in got) > 1000
        with pytest.raises(ValueError):
            fe.observe_features(frames[0][0], frames[0][1][:, :32], frames[0][2], frames[0][3][:, :32])


# ----------------------------------------------------- sharded == unsharded (halo + threshold)

def _run_poses(ctx, frames, first, last, F, P1, P2, keep_from):
    """observe_features over poses [first, last); returns {pose: result} for poses >= keep_from."""
    out = {}
    for p in range(first, last):
        got = ctx.observe_features(p, *frames[p], F, P1, P2, RATIO)
        if p >= keep_from:
            out[p] = got
    return out


def test_sharded_sequence_equals_unsharded():
    """contiguous pose ranges per rank, a read-only halo and no data-path
    collective.  The halo is window + 1 poses: `window` poses whose compacted left frames are
    the resident window of the shard's first pose (src/slam_frontend.cc:424-434), and one pose
    before them whose only purpose is its `mean + 2` threshold (:392-394) - the threshold a
    frame is filtered with depends on nothing but the raw matches of the frame before it."""
    import torch
    from vision_slam_frontend_b200 import sharding
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    W, n_poses, n = 4, 40, 900
    frames = synth.stereo_sequence(n_poses, n, seed=51, overlap=0.85)
    with new_ctx(window=W, max_features=1024) as ctx:
        whole = _run_poses(ctx, frames, 0, n_poses, F, P1, P2, 0)
    world = 2
    ndev = torch.cuda.device_count()
    merged = {}
    for rank in range(world):
        first, last = sharding.pose_range(rank, world, n_poses)
        h0 = sharding.observe_halo_start(first, W)
        with new_ctx(window=W, max_features=1024, device=rank % ndev) as ctx:
            merged.update(_run_poses(ctx, frames, h0, last, F, P1, P2, first))
    assert sorted(merged) == list(range(n_poses))
    n_lists = 0
    for p in range(n_poses):
        a, b = whole[p], merged[p]
        np.testing.assert_array_equal(a["kept_left"], b["kept_left"])
        np.testing.assert_array_equal(a["kept_right"], b["kept_right"])
        assert a["stereo_threshold_next"].view(np.uint32) == b["stereo_threshold_next"].view(np.uint32)
        assert [f for f, _ in a["window"]] == [f for f, _ in b["window"]]
        for (_, ma), (_, mb) in zip(a["window"], b["window"]):
            np.testing.assert_array_equal(ma, mb)
            n_lists += 1
        np.testing.assert_array_equal(a["tri_matches"], b["tri_matches"])
        np.testing.assert_array_equal(a["tri_X4"].view(np.uint32), b["tri_X4"].view(np.uint32))
    assert n_lists == sum(min(p, W) for p in range(n_poses))


# ----------------------------------------- block entry points used by bench.py (frame streams)

@pytest.mark.parametrize("width", [61, 64])
def test_synth_sequence_device_wide_rows_match_numpy_twin(width):
    import torch
    n, poses, stride, seed = 555, 4, 50, 99
    with new_ctx(desc_bytes=width, max_features=1024) as ctx:
        assert ctx.row_bytes == 64
        buf = torch.empty((poses, n, 64), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(buf.data_ptr(), n, 3, poses, stride, seed)
        ctx.synchronize()
        host = buf.cpu().numpy()
    for k in range(poses):
        np.testing.assert_array_equal(host[k][:, :width], synth.synth_pose(n, 3 + k, stride, seed, width))
        assert (host[k][:, width:] == 0).all()


@pytest.mark.parametrize("width", [32, 61])
def test_window_match_block_device(width):
    import torch
    n, W, poses, stride, seed = 1400, 4, 12, 140, 77
    with new_ctx(desc_bytes=width, max_features=2048, window=W) as ctx:
        rb = ctx.row_bytes
        buf = torch.empty((poses, n, rb), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(buf.data_ptr(), n, 0, poses, stride, seed)
        # 11 launches: current frames 4..11 then (wrap) 4, 5, 6
        ctx.window_match_block_device(buf.data_ptr(), n, poses, 0, 11, RATIO)
        got = ctx.fetch_window(W)
    cur = (10 % (poses - W)) + W
    for j in range(W):
        exp = native.get_matches(synth.synth_pose(n, cur - W + j, stride, seed, width),
                                 synth.synth_pose(n, cur, stride, seed, width), RATIO)
        np.testing.assert_array_equal(got[j], exp)
    assert sum(len(m) for m in got) > 1000


@pytest.mark.parametrize("n,W,count,width", [(1400, 4, 1, 32), (1400, 4, 2, 32), (1400, 4, 3, 32), (1400, 4, 6, 32),
                                              (1237, 3, 5, 32), (2900, 2, 4, 32), (700, 10, 9, 32), (130, 38, 3, 32),
                                              (3000, 1, 3, 32), (1400, 4, 6, 61), (1237, 3, 9, 64), (700, 10, 5, 61)])
def test_block_device_launch_variants(n, W, count, width):
    """The poses of a device-resident block are launched in groups, each group as ONE batch of all
    its frame pairs (group * window <= 40 problems): one distance kernel and one finish kernel,
    the next group's frames expanded beside them on a side stream; group 1: a pose is launched alone.  Engine flag
    256 = pose by pose (three kernels), 512 = refine and compaction as separate kernels (four),
    1024 = accepted, no effect any more.  Same lists every way, equal to the oracle's."""
    import torch
    from vision_slam_frontend_b200 import capi
    poses, stride, seed = W + 9, 97, 5
    with new_ctx(desc_bytes=width, max_features=3072, window=W) as ctx:
        buf = torch.empty((poses, n, ctx.row_bytes), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(buf.data_ptr(), n, 0, poses, stride, seed)
        got = {}
        for group, flags, per_pose in ((4, 0, 2), (1, 0, 2), (3, 0, 2), (8, 0, 2), (2, 1024, 2), (1, 1024, 2), (4, 256, 3),
                                       (4, 512, 4), (4, 0, 2)):
            ctx.set_option(capi.OPT_POSE_GROUP, group)
            ctx.set_engine(2, flags)
            before = ctx.launch_count()
            ctx.window_match_block_device(buf.data_ptr(), n, poses, 2, count, RATIO)
            got[len(got)] = ctx.fetch_window(W)
            if width <= 32:
                # + the expansion of the first pose's / group's frames (flags 256 / 512: one per pose, in per_pose)
                if flags & (256 | 512):
                    assert ctx.launch_count() - before == per_pose * count
                else:
                    # per group: the expansion of its frames (side stream), distance, finish
                    g_eff = max(1, min(group, 40 // W))
                    assert ctx.launch_count() - before == 3 * ((count + g_eff - 1) // g_eff)
    cur = ((2 + count - 1) % (poses - W)) + W
    for j in range(W):
        exp = native.get_matches(synth.synth_pose(n, cur - W + j, stride, seed, width),
                                 synth.synth_pose(n, cur, stride, seed, width), RATIO)
        for k in got:
            np.testing.assert_array_equal(got[k][j], exp)


@pytest.mark.parametrize("width", [32, 61])
def test_block_device_large_batches(width):
    """Groups large enough for the finish kernel's 256-query blocks (four poses x ten frames x
    4200 queries = 168 000 >= 150 000) and 128-query blocks (two poses: 84 000 >= 75 000); one pose
    at a time runs the 64-query blocks.  Same lists every way, equal to the oracle's."""
    import torch
    from vision_slam_frontend_b200 import capi
    n, W, count = 4200, 10, 8
    poses, stride, seed = W + 9, 420, 11
    with new_ctx(desc_bytes=width, max_features=4352, window=W) as ctx:
        buf = torch.empty((poses, n, ctx.row_bytes), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(buf.data_ptr(), n, 0, poses, stride, seed)
        ctx.set_engine(2, 0)
        got = []
        for group in (4, 2, 1):
            ctx.set_option(capi.OPT_POSE_GROUP, group)
            ctx.window_match_block_device(buf.data_ptr(), n, poses, 1, count, RATIO)
            got.append(ctx.fetch_window(W))
    cur = ((1 + count - 1) % (poses - W)) + W
    total = 0
    for j in range(W):
        exp = native.get_matches(synth.synth_pose(n, cur - W + j, stride, seed, width),
                                 synth.synth_pose(n, cur, stride, seed, width), RATIO)
        total += len(exp)
        for g in got:
            np.testing.assert_array_equal(g[j], exp)
    assert total > 5000


@pytest.mark.parametrize("width", [32, 61])
def test_tensor_launch_replayed_from_a_cuda_graph(width):
    """The finish kernel's ticket counter and look-back epoch live on the device and are advanced by
    the launch itself, so a captured launch sequence can be replayed (bench.py's C2 leg does) - also
    on new data in the same buffers."""
    import ctypes as C
    import torch
    n, W = 1500, 2
    with new_ctx(desc_bytes=width, max_features=2048, window=W) as ctx:
        stream = torch.cuda.Stream()
        ctx.set_stream(stream.cuda_stream)
        ctx.set_engine(2, 0)
        rb = ctx.row_bytes

        def padded(a):
            out = np.zeros((a.shape[0], rb), np.uint8)
            out[:, :width] = a
            return out
        q = [torch.zeros((n, rb), dtype=torch.uint8, device="cuda") for _ in range(W)]
        t = torch.zeros((n, rb), dtype=torch.uint8, device="cuda")
        qp = (C.c_void_p * W)(*[x.data_ptr() for x in q])
        nn = (C.c_int * W)(*([n] * W))

        def launch():
            assert ctx._L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(t.data_ptr()), n, RATIO) == 0

        def load(seed):
            frames = [synth.synth_pose(n, p, 150, seed, width) for p in range(W + 1)]
            for j in range(W):
                q[j].copy_(torch.from_numpy(padded(frames[j])))
            t.copy_(torch.from_numpy(padded(frames[W])))
            torch.cuda.synchronize()
            return frames
        load(1)
        with torch.cuda.stream(stream):
            for _ in range(3):
                launch()
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                launch()
        for seed in (2, 3, 4, 5):
            frames = load(seed)
            g.replay()
            torch.cuda.synchronize()
            got = ctx.fetch_window(W)
            for j in range(W):
                np.testing.assert_array_equal(got[j], native.get_matches(frames[j], frames[W], RATIO))
        launch()     # and an ordinary launch after the replays
        got = ctx.fetch_window(W)
        for j in range(W):
            np.testing.assert_array_equal(got[j], native.get_matches(frames[j], frames[W], RATIO))


@pytest.mark.parametrize("sort_mode,width,lag,shape", [
    (0, 32, 12, None), (1, 32, 12, None), (2, 32, 12, None), (1, 61, 12, None), (2, 61, 9, None), (1, 32, 5, None),
    (2, 32, 1, None),
    # batches of four frames x ten past frames that are large enough for the finish kernel's 256- and
    # 128-query blocks (>= 150 000 / 75 000 queries per batch), both widths
    (1, 32, 12, (4200, 10, 15, 9)), (0, 61, 12, (4200, 10, 15, 9)), (1, 32, 12, (2100, 10, 15, 9)),
    (2, 61, 12, (2100, 10, 15, 9))])
def test_window_run_sequence_equals_per_frame_calls(sort_mode, width, lag, shape):
    """vsf_window_run_sequence (groups of frames uploaded, then launched as one batch) returns per
    frame what the oracle's GetMatches + sort + cut gives for the window of that moment."""
    import torch
    from vision_slam_frontend_b200 import capi
    n, W, n_pool, count = shape or (1100, 3, 9, 14)
    rb = 32 if width <= 32 else 64
    frames = [synth.synth_pose(n, p, max(1, n // 10), 3, width) for p in range(n_pool)]
    padded = np.zeros((n_pool, n, rb), np.uint8)
    for p in range(n_pool):
        padded[p, :, :width] = frames[p]
    pool = torch.from_numpy(padded).pin_memory()
    hp = pool.numpy()
    with new_ctx(desc_bytes=width, max_features=max(2048, n), window=W) as ctx:
        ctx.set_engine(2, 0)
        for p in range(W):
            ctx.window_push(1000 + p, frames[p])
        out = np.zeros((count, W, n), capi.FEATURE_MATCH_DTYPE)
        counts = np.zeros((count, W), np.int32)
        h2d, d2h = ctx.window_run_sequence(hp, W, count, RATIO, float(BP), sort_mode, lag, out, counts)
        assert h2d == count * n * rb and d2h > 0 and ctx.window_in_flight() == 0
    live = [frames[p] for p in range(W)]
    for k in range(count):
        D = frames[(W + k) % n_pool]
        for j, past in enumerate(live):
            m = native.get_matches(past, D, RATIO)
            keep = restate.num_good_matches(len(m), BP)
            order = restate.sort_order_stdsort(m) if sort_mode >= 1 else restate.sort_order_stable(m)
            exp = m[order][:keep]
            assert counts[k, j] == keep
            np.testing.assert_array_equal(out[k, j, :keep]["feature_idx_initial"], exp["queryIdx"].astype(np.uint64))
            np.testing.assert_array_equal(out[k, j, :keep]["feature_idx_current"], exp["trainIdx"].astype(np.uint64))
        live.pop(0)
        live.append(D)


# ------------------------------------------- the whole frame, pipelined (vsf_observe_submit / collect)

@pytest.mark.parametrize("depth", [1, 4])
def test_observe_submit_collect_against_frontend_oracle(depth):
    """Up to VSF_OBSERVE_DEPTH whole frames in flight: everything a frame needs from its
    predecessors (compacted frames, their row counts, the adaptive threshold) stays on the
    device.  Every collected frame equals the oracle's, N1 included."""
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    K = synth.KITTI_K.astype(np.float32)
    dist = np.array([-0.153137, 0.075666, -0.000227, -0.000320, 0.01], np.float32)
    W = 3
    frames = synth.stereo_sequence(9, 1800, seed=61)
    frames[4] = tuple(a[:700] for a in frames[4])          # a smaller frame in the middle of the stream
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order="stable")
    expected = []
    with new_ctx(window=W, max_features=2048) as ctx:
        def check_one():
            got = ctx.observe_collect()
            p, past, r, thr = expected.pop(0)
            assert got["frame_id"] == p
            sm = r.stereo_matches
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][r.stereo_keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][r.stereo_keep])
            assert same_f32(got["stereo_threshold_next"], thr)
            assert [fid for fid, _ in got["window"]] == [f.frame_ID for f in past]
            for (fid, m), pf in zip(got["window"], past):
                np.testing.assert_array_equal(m, native.get_matches(pf.descriptors, r.left.descriptors, RATIO))
            tm = native.get_matches(r.right.descriptors, r.left.descriptors, RATIO)
            np.testing.assert_array_equal(got["tri_matches"], tm)
            order = restate.sort_order_stable(tm)
            pts = restate.dehomogenize(got["tri_X4"].T)[order]
            rel = np.abs(pts - r.points).max(1) / np.abs(r.points).max(1)
            assert rel.max() < 1e-4, rel.max()
            np.testing.assert_allclose(got["xy_undist"], restate.undistort_points(r.features_pixel, K, dist), atol=2e-3)

        for p, (kl, dl, kr, dr) in enumerate(frames):
            past = list(fo.frame_list)
            r = fo.observe_features(kl, dl, kr, dr)
            expected.append((p, past, r, fo.stereo_ambig_constraint))
            ctx.observe_submit(p, kl, dl, kr, dr, F, P1, P2, K, dist, RATIO)
            assert ctx.observe_in_flight() == len(expected)
            while ctx.observe_in_flight() >= depth:
                check_one()
        while expected:
            check_one()
        assert ctx.window_size() == W
        assert same_f32(ctx.get_stereo_threshold(), fo.stereo_ambig_constraint)


@pytest.mark.parametrize("width", [32, 61])
def test_frontend_pipelined_equals_blocking(width):
    """Frontend::SubmitFeatures / CollectFeatures with several frames pending produce the same
    SLAMProblem bytes as ObserveFeatures frame by frame."""
    from vision_slam_frontend_b200.frontend import Frontend, synthetic_rig, synth_frame
    rig = synthetic_rig()
    n, W, poses = 1200, 3, 9
    seq = [synth_frame(n, width, 7, p) for p in range(poses)]

    def run(pipelined):
        with Frontend(max_features=n, desc_bytes=width, frame_life=W, P_left=rig["P_left"], P_right=rig["P_right"],
                      fundamental=rig["fundamental"], K_left=rig["K_left"], dist_left=rig["dist_left"]) as fe:
            for p, (kl, dl, kr, dr, t, q, ts) in enumerate(seq):
                fe.observe_odometry(t, q, ts)
                if pipelined:
                    assert fe.submit_features(kl, dl, kr, dr, ts)
                    while fe.in_flight >= 3:
                        assert fe.collect_features()
                else:
                    assert fe.observe_features(kl, dl, kr, dr, ts)
            while fe.collect_features():
                pass
            assert fe.num_poses == poses
            return fe.serialize_problem()

    a, b = run(False), run(True)
    assert len(a) > 100000 and a == b


def test_synthetic_sequence_against_frontend_oracle():
    """The C++ chain synthetic source -> Frontend (pipelined) -> SLAMProblem message against the
    FrontendOracle fed with the same frames: every vision factor and node."""
    from vision_slam_frontend_b200.frontend import (parse_slam_problem, run_synthetic_sequence, synth_frame,
                                                    synthetic_rig)
    rig = synthetic_rig()
    n, W, poses, width, seed = 1500, 4, 9, 61, 5
    msg = parse_slam_problem(run_synthetic_sequence(n, width, W, poses, world=1, in_flight=3, seed=seed))
    fo = restate.FrontendOracle(rig["P_left"], rig["P_right"], rig["fundamental"], frame_life=W, order="stdsort")
    results = []
    for p in range(poses):
        kl, dl, kr, dr, t, q, ts = synth_frame(n, width, seed, p)
        results.append((fo.observe_features(kl, dl, kr, dr), t, ts))
    assert len(msg["vision_factors"]) == len(fo.vision_factors) == sum(min(p, W) for p in range(poses))
    for (a, b, pairs), e in zip(msg["vision_factors"], fo.vision_factors):
        assert (a, b) == (e.pose_idx_initial, e.pose_idx_current)
        np.testing.assert_array_equal(pairs, e.feature_matches)
    assert sum(len(p) for _, _, p in msg["vision_factors"]) > 1000
    assert len(msg["nodes"]) == poses and len(msg["odometry_factors"]) == poses - 1
    for p, (r, t, ts) in enumerate(results):
        node = msg["nodes"][p]
        assert node["id"] == p and node["timestamp"] == ts and len(node["features"]) == len(r.features_pixel)
        pix = np.array([f[1][:2] for f in node["features"]], np.float32)
        np.testing.assert_allclose(pix, restate.undistort_points(r.features_pixel, rig["K_left"], rig["dist_left"]),
                                   atol=2e-3)
        p3 = np.array([f[2] for f in node["features"]], np.float32)
        np.testing.assert_array_equal(np.isnan(p3), np.isnan(r.features_point3d))
        ok = ~np.isnan(p3).any(1)
        rel = np.abs(p3[ok] - r.features_point3d[ok]).max(1) / np.abs(r.features_point3d[ok]).max(1)
        assert rel.max() < 1e-4


@pytest.mark.parametrize("width", [32, 61])
def test_synthetic_sequence_sharded_bytes_equal_unsharded(width):
    """SURVEY 8(e) at the level the reference writes its output (src/slam_frontend_main.cc:369-374):
    the SLAMProblem message of a sequence run as 2 and 3 pose-range shards (each with its
    (window + 1)-frame halo, pipelined) is byte-identical to the unsharded one."""
    from vision_slam_frontend_b200.frontend import run_synthetic_sequence
    n, W, poses = 1000, 3, 23
    whole = run_synthetic_sequence(n, width, W, poses, world=1, in_flight=1)
    assert len(whole) > 300000
    for world, depth in ((2, 3), (3, 4)):
        assert run_synthetic_sequence(n, width, W, poses, world=world, in_flight=depth) == whole


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_sequence_driver_two_ranks_over_nccl(tmp_path):
    """vsf_sequence_driver as 2 processes on 2 GPUs: pieces gathered to rank 0 with
    vsf_nccl_gather_bytes; the file equals the one a single process writes."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "vision_slam_frontend_b200", "vsf_sequence_driver")
    common = ["--poses", "30", "--features", "1200", "--window", "4", "--desc-bytes", "61"]
    one = str(tmp_path / "one.bin")
    subprocess.run([exe, *common, "--world", "1", "--rank", "0", "--device", "0", "--out", one], check=True, timeout=300)
    two = str(tmp_path / "two.bin")
    rdv = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, *common, "--world", "2", "--rank", str(r), "--device", str(r), "--out", two,
                               "--rendezvous", rdv]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    a, b = open(one, "rb").read(), open(two, "rb").read()
    assert len(a) > 100000 and a == b


def test_gather_matches_two_ranks_over_nccl(tmp_path):
    """vsf_gather_matches under torchrun on 2 GPUs: every rank ends up with every rank's device
    match lists, equal to the oracle's for that rank's pose."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(root, "tools", "nccl_gather_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("gather ok") == 2


# ------------------------------------ sort_mode 2: libstdc++'s std::sort order, replayed on the device

def _dmatch_list(dist):
    import vision_slam_frontend_b200 as vsf
    m = np.zeros(len(dist), vsf.DMATCH_DTYPE)
    m["queryIdx"] = np.arange(len(dist)) * 3 + 1            # not the position: the kernel must carry the record
    m["trainIdx"] = (np.arange(len(dist)) * 7 + 5) % 100003
    m["distance"] = np.asarray(dist, np.float32)
    return m


@pytest.mark.parametrize("best_percent", [0.3, 1.0, 0.05, 0.77])
def test_device_exact_sort_equals_std_sort(best_percent):
    """The corpus of tests/test_exact_sort.py (sizes around the 16-element threshold, powers of
    two, C4 / C5 list lengths; from all-equal to 257 distinct distances) through the device
    replay: the kept prefix is the real std::sort's (oracle/stdsort_oracle.cc), ties included."""
    rng = np.random.default_rng(int(best_percent * 1000) + 11)
    sizes = list(range(0, 70)) + [100, 255, 256, 257, 500, 1000, 2500, 4500, 5000, 12000, 20000]
    with new_ctx(max_features=20000, window=2) as ctx:
        for n in sizes:
            for spread in (1, 2, 7, 40, 257):
                dist = rng.integers(0, spread, size=n)
                m = _dmatch_list(dist)
                keep = restate.num_good_matches(n, np.float32(best_percent))
                exp = m[restate.sort_order_stdsort(m)][:keep]
                got = ctx.debug_sort_device(m, best_percent, exact=True)
                assert len(got) == keep, (n, spread)
                np.testing.assert_array_equal(got[:, 0], exp["queryIdx"].astype(np.uint64), err_msg=str((n, spread)))
                np.testing.assert_array_equal(got[:, 1], exp["trainIdx"].astype(np.uint64))
                # and the stable mode on the same list
                exp0 = m[restate.sort_order_stable(m)][:keep]
                got0 = ctx.debug_sort_device(m, best_percent, exact=False)
                np.testing.assert_array_equal(got0[:, 0], exp0["queryIdx"].astype(np.uint64))


def test_device_exact_sort_sorted_and_adversarial_inputs():
    rng = np.random.default_rng(4)
    cases = [np.arange(3000) // 12, (np.arange(3000) // 12)[::-1], np.tile([5, 1, 9, 1, 5], 700),
             np.concatenate([np.full(2000, 30), rng.integers(0, 30, 1500)]),
             np.where(np.arange(4096) % 2 == 0, 7, rng.integers(0, 200, 4096))]
    with new_ctx(max_features=8192, window=2) as ctx:
        for dist in cases:
            m = _dmatch_list(dist)
            for bp in (0.3, 1.0):
                keep = restate.num_good_matches(len(m), np.float32(bp))
                exp = m[restate.sort_order_stdsort(m)][:keep]
                got = ctx.debug_sort_device(m, bp, exact=True)
                np.testing.assert_array_equal(got[:, 0], exp["queryIdx"].astype(np.uint64))


@pytest.mark.parametrize("depth", [0, 1, 2, 4])
def test_device_exact_sort_heapsort_fallback(depth):
    """Introsort's depth limit forced low on both sides (VSF_OPT_DEBUG_SORT_DEPTH /
    vsf_debug_sort_prefix_depth): the ranges that hit it go through libstdc++'s heapsort, whose
    sift order the kernel replays too."""
    import vision_slam_frontend_b200 as vsf
    from vision_slam_frontend_b200 import capi
    rng = np.random.default_rng(depth)
    lib = vsf.load_library()
    with new_ctx(max_features=8192, window=2) as ctx:
        ctx.set_option(capi.OPT_DEBUG_SORT_DEPTH, depth)
        for n in (17, 40, 100, 1000, 4500):
            for spread in (2, 40, 300):
                dist = rng.integers(0, spread, size=n)
                m = _dmatch_list(dist)
                keep = int(np.float32(n) * np.float32(0.3))
                keys = np.ascontiguousarray((dist.astype(np.uint32) << 22) | np.arange(n, dtype=np.uint32))
                assert lib.vsf_debug_sort_prefix_depth(keys.ctypes.data, n, keep, depth) == 0
                exp = m[(keys[:keep] & np.uint32((1 << 22) - 1)).astype(np.int64)]
                got = ctx.debug_sort_device(m, 0.3, exact=True)
                np.testing.assert_array_equal(got[:, 0], exp["queryIdx"].astype(np.uint64), err_msg=str((n, spread)))
        ctx.set_option(capi.OPT_DEBUG_SORT_DEPTH, -1)
