"""Parity of the CUDA path (through the C ABI) with the oracle and the OpenCV
golden vectors.  Integer / index work is compared bit-exactly; triangulated
points to the 1e-4 relative tolerance BASELINE.json's north_star states."""
import glob
import os

import numpy as np
import pytest

import synth
from oracle import native, restate

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATIO = restate.NN_MATCH_RATIO
TRI_RTOL = 1e-4      # north_star: "Triangulated points must agree within 1e-4 relative error"


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def new_ctx(**kw):
    import vision_slam_frontend_b200 as vsf
    args = dict(device=0, max_features=4096, desc_bytes=32, window=10)
    args.update(kw)
    return vsf.Context(**args)


# ------------------------------------------------------------------ a1: kNN

@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4]
                                        for p in glob.glob(os.path.join(GOLD, "knn_*.npz"))))
def test_knn2_against_opencv_golden(name):
    g = load(name)
    with new_ctx(desc_bytes=g["Q"].shape[1]) as ctx:
        idx, dist = ctx.knn2(g["Q"], g["T"])
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(dist, g["dist"])


@pytest.mark.parametrize("mode", [-1, 0, 2, 3])
@pytest.mark.parametrize("R", [0, 1, 2, 4])
@pytest.mark.parametrize("split,variant", [(0, -1), (1, 0), (5, 1), (32, 2), (3, 3), (1, 1), (2, 2)])
def test_knn2_c2_all_kernel_variants(vsf_ctx, mode, R, split, variant):
    Q, T = synth.descriptor_pair(2000, 2000, seed=0)
    ei, ed = native.knn2_hamming(Q, T)
    vsf_ctx.set_tuning(mode, split, R, variant)
    try:
        idx, dist = vsf_ctx.knn2(Q, T)
    finally:
        vsf_ctx.set_tuning()
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)


@pytest.mark.parametrize("split", [0, 1, 7])
def test_knn2_ties_lowest_train_index(vsf_ctx, split):
    Q, T = synth.tie_pair(1500, 2300, seed=3)
    ei, ed = native.knn2_hamming(Q, T)
    vsf_ctx.set_tuning(-1, split, 0, split % 4)
    try:
        idx, dist = vsf_ctx.knn2(Q, T)
    finally:
        vsf_ctx.set_tuning()
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)
    assert (idx[:, 0] < idx[:, 1]).sum() > 100     # ties did occur


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (3, 2), (31, 33), (33, 31), (257, 129),
                                   (1000, 17), (17, 1000), (129, 4097)])
def test_knn2_ragged_shapes(vsf_ctx, nq, nt):
    Q, T = synth.descriptor_pair(nq, nt, seed=nq * 7 + nt)
    ei, ed = native.knn2_hamming(Q, T)
    idx, dist = vsf_ctx.knn2(Q, T)
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)


def test_knn2_empty_inputs(vsf_ctx):
    Q, T = synth.descriptor_pair(40, 30, seed=1)
    idx, dist = vsf_ctx.knn2(Q[:0], T)
    assert idx.shape == (0, 2)
    idx, dist = vsf_ctx.knn2(Q, T[:0])
    assert (idx == -1).all() and (dist == -1).all()
    assert len(vsf_ctx.get_matches(Q, T[:0], RATIO)) == 0
    assert len(vsf_ctx.get_matches(Q, T[:1], RATIO)) == 0      # quirk Q6
    assert len(vsf_ctx.get_matches(Q[:0], T, RATIO)) == 0


def test_knn2_strided_rows_and_64_byte_descriptors():
    Q, T = synth.descriptor_pair(700, 900, width=64, seed=8)
    Qs = np.zeros((700, 80), np.uint8)
    Qs[:, :64] = Q
    ei, ed = native.knn2_hamming(Q, T)
    with new_ctx(desc_bytes=64) as ctx:
        idx, dist = ctx.knn2(Qs[:, :64], T)       # cv::Mat with step 80
        np.testing.assert_array_equal(idx, ei)
        np.testing.assert_array_equal(dist, ed)
        for mode, var in ((0, 1), (2, 2), (3, 3)):
            ctx.set_tuning(mode, 3, 1, var)
            idx, dist = ctx.knn2(Q, T)
            np.testing.assert_array_equal(idx, ei)
            np.testing.assert_array_equal(dist, ed)


def test_knn2_akaze_61_byte_descriptors():
    Q, T = synth.descriptor_pair(500, 600, width=61, seed=18)
    ei, ed = native.knn2_hamming(Q, T)
    with new_ctx(desc_bytes=61) as ctx:
        idx, dist = ctx.knn2(Q, T)
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)


def test_knn2_full_size_properties(vsf_ctx):
    """C5's frame size (20000 rows): checked against the C oracle and through
    size-independent properties (self-match, permutation equivariance)."""
    Q, T = synth.descriptor_pair(20000, 20000, seed=12, planted=0.5)
    idx, dist = vsf_ctx.knn2(Q, T)
    ei, ed = native.knn2_hamming(Q, T)
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)
    # every row is its own nearest neighbour at distance 0 (rows are unique)
    si, sd = vsf_ctx.knn2(T, T)
    np.testing.assert_array_equal(si[:, 0], np.arange(20000))
    assert (sd[:, 0] == 0).all() and (sd[:, 1] > 0).all()
    # permuting the train rows permutes the indices and keeps the distances
    rng = np.random.default_rng(5)
    p = rng.permutation(20000)
    pi, pd = vsf_ctx.knn2(Q[:4096], T[p])
    np.testing.assert_array_equal(pd, dist[:4096])
    strict = dist[:4096, 0] < dist[:4096, 1]
    np.testing.assert_array_equal(p[pi[strict, 0]], idx[:4096][strict, 0])


# ------------------------------------------------------------- a2: GetMatches

@pytest.mark.parametrize("ratio", [RATIO, 0.8, 1.0, 0.0, 0.3333333333333333])
def test_get_matches_against_oracle(vsf_ctx, ratio):
    Q, T = synth.descriptor_pair(2000, 2000, seed=2, max_flips=60)
    exp = native.get_matches(Q, T, ratio)
    got = vsf_ctx.get_matches(Q, T, ratio)
    np.testing.assert_array_equal(got, exp)
    np.testing.assert_array_equal(got, restate.get_matches(Q, T, ratio))


# ---------------------------------------------------- a4: the sliding window

def test_window_match_sequence_against_oracle():
    n, W = 1500, 4
    frames = [synth.synth_pose(n - 37 * (p % 3), p, 150, 99) for p in range(8)]
    with new_ctx(window=W) as ctx:
        live = []
        for p, D in enumerate(frames):
            got = ctx.window_match(D, RATIO)
            assert [fid for fid, _ in got] == [fid for fid, _ in live]
            for (fid, m), (_, past) in zip(got, live):
                np.testing.assert_array_equal(m, native.get_matches(past, D, RATIO))
            if p % 2 == 0:
                ctx.window_commit(p, len(D))          # zero-copy push of the frame just matched
            else:
                ctx.window_push(p, D)
            if len(live) >= W:                        # src/slam_frontend.cc:467-470
                live.pop(0)
            live.append((p, D))
            assert ctx.window_size() == len(live)
        assert sum(len(m) for _, m in got) > 1000


@pytest.mark.parametrize("sort_mode", [0, 1, 2])
def test_window_feature_matches(sort_mode):
    n, W = 1200, 3
    frames = [synth.synth_pose(n, p, 120, 5) for p in range(5)]
    bp = restate.BEST_PERCENT
    with new_ctx(window=W) as ctx:
        live = []
        for p, D in enumerate(frames):
            got = ctx.window_feature_matches(D, RATIO, float(bp), sort_mode)
            assert len(got) == len(live)
            for (fid, pairs), (pid, past) in zip(got, live):
                assert fid == pid
                m = native.get_matches(past, D, RATIO)
                keep = restate.num_good_matches(len(m), bp)
                assert len(pairs) == keep
                order = restate.sort_order_stdsort(m) if sort_mode >= 1 else restate.sort_order_stable(m)
                exp = m[order][:keep]
                np.testing.assert_array_equal(pairs[:, 0], exp["queryIdx"].astype(np.uint64))
                np.testing.assert_array_equal(pairs[:, 1], exp["trainIdx"].astype(np.uint64))
            ctx.window_commit(p, len(D))
            if len(live) >= W:
                live.pop(0)
            live.append((p, D))


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("engine", [1, 2])
@pytest.mark.parametrize("sort_mode", [0, 1, 2])
def test_window_pipelined_submit_collect(sort_mode, engine, pinned):
    """vsf_window_submit / vsf_window_collect with several frames in flight return, frame by
    frame, what the reference's loop (src/slam_frontend.cc:424-434, :467-470) produces."""
    n, W = 1300, 3
    frames = [synth.synth_pose(n - 53 * (p % 4), p, 130, 17) for p in range(13)]
    frames[5] = frames[5][:0]                       # an empty frame in the middle of the stream
    if pinned:                                      # VSF_SUBMIT_PINNED_DESC: page-locked rows, read in place
        import torch
        keep_alive = [torch.from_numpy(np.ascontiguousarray(D)).pin_memory() for D in frames]
        frames = [t.numpy() for t in keep_alive]
    bp = restate.BEST_PERCENT
    import vision_slam_frontend_b200 as vsf
    depth = vsf.PIPELINE_DEPTH                      # VSF_PIPELINE_DEPTH
    with new_ctx(window=W) as ctx:
        ctx.set_engine(engine, 0)
        with pytest.raises(vsf.VsfError) as e:
            ctx.window_collect()
        assert e.value.code == 4                    # VSF_ERR_STATE: nothing in flight
        live, expected, submitted, collected = [], [], 0, 0

        def check_one():
            fid, got = ctx.window_collect()
            efid, elive, D = expected.pop(0)
            assert fid == efid and len(got) == len(elive)
            for (pfid, pairs), (pid, past) in zip(got, elive):
                assert pfid == pid
                m = native.get_matches(past, D, RATIO)
                keep = restate.num_good_matches(len(m), bp)
                assert len(pairs) == keep
                order = restate.sort_order_stdsort(m) if sort_mode >= 1 else restate.sort_order_stable(m)
                exp = m[order][:keep]
                np.testing.assert_array_equal(pairs[:, 0], exp["queryIdx"].astype(np.uint64))
                np.testing.assert_array_equal(pairs[:, 1], exp["trainIdx"].astype(np.uint64))

        for p, D in enumerate(frames):
            if ctx.window_in_flight() == depth:
                with pytest.raises(vsf.VsfError) as e:
                    ctx.window_submit(100 + p, D, RATIO, float(bp), sort_mode, pinned)
                assert e.value.code == 4            # full: collect first
                check_one()
                collected += 1
            ctx.window_submit(100 + p, D, RATIO, float(bp), sort_mode, pinned)
            submitted += 1
            expected.append((100 + p, list(live), D))
            if len(live) >= W:
                live.pop(0)
            live.append((100 + p, D))
            assert ctx.window_size() == len(live)
            assert ctx.window_in_flight() == submitted - collected
        while expected:
            check_one()
        assert ctx.window_in_flight() == 0


def test_window_c4_full_size(vsf_ctx):
    """BASELINE config 4's per-pose shape: 5000 features against 10 prior frames."""
    n, W = 5000, 10
    vsf_ctx.window_clear()
    frames = [synth.synth_pose(n, p, 500, 7) for p in range(W + 1)]
    for p in range(W):
        vsf_ctx.window_push(p, frames[p])
    got = vsf_ctx.window_match(frames[W], RATIO)
    assert len(got) == W
    total = 0
    for j, (fid, m) in enumerate(got):
        assert fid == j
        np.testing.assert_array_equal(m, native.get_matches(frames[j], frames[W], RATIO))
        total += len(m)
    assert total > 10000
    vsf_ctx.window_clear()


# --------------------------------------------------- a5: stereo filter

def test_stereo_filter_sequence_against_oracle():
    F = synth.kitti_fundamental()
    frames = synth.stereo_sequence(4, 2000, seed=3)
    thresh = restate.STEREO_AMBIG_INIT
    with new_ctx() as ctx:
        for kl, dl, kr, dr in frames:
            got = ctx.stereo_filter(kl, dl, kr, dr, F, RATIO)
            sm = native.get_matches(dl, dr, RATIO)
            L, R, thresh_next, c, keep = restate.remove_ambig_stereo(
                restate.Frame(kl, dl, 0), restate.Frame(kr, dr, 0), sm, F, thresh)
            np.testing.assert_array_equal(got["stereo_matches"], sm)
            np.testing.assert_array_equal(got["residuals"].view(np.uint32), c.view(np.uint32))
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][keep])
            assert ctx.get_stereo_threshold().view(np.uint32) == np.float32(thresh_next).view(np.uint32)
            thresh = thresh_next
        assert keep.sum() > 1000 and (~keep).sum() > 20


def test_stereo_threshold_nan_when_no_matches():
    F = synth.kitti_fundamental()
    rng = np.random.default_rng(0)
    kl = synth.make_keypoints(rng.uniform(0, 300, (50, 2)))
    d1 = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    with new_ctx() as ctx:
        got = ctx.stereo_filter(kl, d1, kl, d2, F, RATIO)     # uniform rows never pass the ratio test
        assert len(got["stereo_matches"]) == 0 and len(got["kept_left"]) == 0
        assert np.isnan(ctx.get_stereo_threshold())           # 0/0, like the reference (quirk Q4)
        ctx.set_stereo_threshold(10000.0)
        assert ctx.get_stereo_threshold() == np.float32(10000.0)


# ------------------------------------------------ a6: triangulation

@pytest.mark.parametrize("name", ["triangulate_kitti", "triangulate_pointgrey"])
def test_triangulate_against_opencv_golden(vsf_ctx, name):
    g = load(name)
    X4 = vsf_ctx.triangulate(g["P1"], g["P2"], g["x1"], g["x2"])
    assert X4.shape == g["X4"].shape and X4.dtype == np.float32
    a, b = restate.dehomogenize(X4), restate.dehomogenize(g["X4"])
    rel = np.abs(a - b).max(1) / np.abs(b).max(1)
    assert rel.max() < TRI_RTOL, rel.max()
    np.testing.assert_allclose(np.linalg.norm(X4, axis=0), 1.0, atol=1e-6)


def test_triangulate_c3_size_against_oracle(vsf_ctx):
    P1, P2 = synth.kitti_projections()
    kl, dl, kr, dr, X, perm = synth.stereo_frame(2000, seed=1, bad_geometry=0.0)
    ok = perm >= 0
    x1 = np.stack([kl["x"][ok], kl["y"][ok]], 1)
    x2 = np.stack([kr["x"][perm[ok]], kr["y"][perm[ok]]], 1)
    a = restate.dehomogenize(vsf_ctx.triangulate(P1, P2, x1, x2))
    b = restate.dehomogenize(restate.triangulate_points(P1, P2, x1, x2))
    rel = np.abs(a - b).max(1) / np.abs(b).max(1)
    assert rel.max() < TRI_RTOL, rel.max()
    # and it actually recovers the scene (0.5 px noise): median depth error below 5 %
    assert np.median(np.abs(a[:, 2] - X[ok, 2]) / X[ok, 2]) < 0.05


# ------------------------------- the fused ObserveImage matching path

def test_observe_features_sequence_against_frontend_oracle():
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    W = 3
    frames = synth.stereo_sequence(6, 2000, seed=9)
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order="stable")
    with new_ctx(window=W) as ctx:
        for p, (kl, dl, kr, dr) in enumerate(frames):
            past = list(fo.frame_list)
            r = fo.observe_features(kl, dl, kr, dr)
            got = ctx.observe_features(p, kl, dl, kr, dr, F, P1, P2, RATIO)
            sm = r.stereo_matches
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][r.stereo_keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][r.stereo_keep])
            assert got["stereo_threshold_next"].view(np.uint32) == \
                np.float32(fo.stereo_ambig_constraint).view(np.uint32)
            # window stage: query-ordered GetMatches per resident frame, oldest first
            assert [fid for fid, _ in got["window"]] == [f.frame_ID for f in past]
            for (fid, m), pf in zip(got["window"], past):
                np.testing.assert_array_equal(m, native.get_matches(pf.descriptors, r.left.descriptors, RATIO))
            # triangulation stage: R'->L' matches in query order + their points
            tm = native.get_matches(r.right.descriptors, r.left.descriptors, RATIO)
            np.testing.assert_array_equal(got["tri_matches"], tm)
            order = restate.sort_order_stable(tm)
            np.testing.assert_array_equal(tm[order], r.tri_matches)
            pts = restate.dehomogenize(got["tri_X4"].T)[order]
            rel = np.abs(pts - r.points).max(1) / np.abs(r.points).max(1)
            assert rel.max() < TRI_RTOL, rel.max()
            assert ctx.window_size() == len(fo.frame_list)
        assert len(r.points) > 1000


# ------------------------------------ bench / test frame source

def test_synth_sequence_device_matches_numpy_twin(vsf_ctx):
    import torch
    n, poses, stride, seed = 777, 5, 70, 1234
    buf = torch.empty((poses, n, 32), dtype=torch.uint8, device="cuda")
    vsf_ctx.synth_sequence_device(buf.data_ptr(), n, 2, poses, stride, seed)
    vsf_ctx.synchronize()
    host = buf.cpu().numpy()
    for k in range(poses):
        np.testing.assert_array_equal(host[k], synth.synth_pose(n, 2 + k, stride, seed))


def test_window_match_device_entry_point(vsf_ctx):
    import torch
    n, W, stride, seed = 3000, 4, 300, 77
    buf = torch.empty((W + 1, n, 32), dtype=torch.uint8, device="cuda")
    vsf_ctx.synth_sequence_device(buf.data_ptr(), n, 0, W + 1, stride, seed)
    base = buf.data_ptr()
    vsf_ctx.window_match_device([base + j * n * 32 for j in range(W)], [n] * W, base + W * n * 32, n, RATIO)
    got = vsf_ctx.fetch_window(W)
    host = buf.cpu().numpy()
    for j in range(W):
        np.testing.assert_array_equal(got[j], native.get_matches(host[j], host[W], RATIO))
    assert len(got[W - 1]) > 2000


def test_probe_pipe_reports_plausible_rates(vsf_ctx):
    popc = vsf_ctx.probe_pipe(0, 2048)
    lop3 = vsf_ctx.probe_pipe(1, 2048)
    assert 1e11 < popc < 1e14 and 1e11 < lop3 < 1e14


def test_undistort_points_device_matches_opencv_golden(vsf_ctx):
    """N1: cv::undistortPoints(pts, K, dist, R = {}, P = K) (src/slam_frontend.cc:323-351) on the
    device: the OpenCV golden vector and the oracle restatement.  Floating point: fp64 internals,
    float output; tolerance 1e-4 relative (north_star), measured at float rounding."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort_pointgrey.npz"))
    got = vsf_ctx.undistort_points(g["K"], g["dist"], g["px"])
    exp = np.asarray(g["out"], np.float32).reshape(-1, 2)
    rel = np.abs(got - exp).max() / np.abs(exp).max()
    assert rel < 1e-4, rel
    np.testing.assert_allclose(got, restate.undistort_points(g["px"], g["K"], g["dist"]), rtol=0, atol=2e-3)
    # a wider field of points, strong distortion, empty input
    rng = np.random.default_rng(2)
    px = (rng.random((5000, 2)) * np.float32([1280, 1024])).astype(np.float32)
    K = np.float32([[700, 0, 640], [0, 705, 512], [0, 0, 1]])
    dist = np.float32([-0.31, 0.12, 1e-3, -2e-3, -0.02])
    got = vsf_ctx.undistort_points(K, dist, px)
    exp = restate.undistort_points(px, K, dist)
    assert np.abs(got - exp).max() / np.abs(exp).max() < 1e-4
    assert vsf_ctx.undistort_points(K, dist, np.zeros((0, 2), np.float32)).shape == (0, 2)
