"""Parity of the tensor-core kNN engine (tcgen05 int8 / e4m3 MMA over +-1 expanded
descriptor bits, csrc/knn2_tc_kernel.cu) with the oracle and the OpenCV golden
vectors, through the C ABI.  Everything here is integer / index work: bit-exact."""
import glob
import os

import numpy as np
import pytest

import synth
from oracle import native, restate

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATIO = restate.NN_MATCH_RATIO
ENGINES = [2, 3]          # tensor cores: int8 operands, e4m3 operands


@pytest.fixture(params=ENGINES, ids=["int8", "e4m3"])
def tc_ctx(request, vsf_ctx):
    vsf_ctx.set_engine(request.param, 0)
    vsf_ctx.set_tuning()
    yield vsf_ctx
    vsf_ctx.set_engine(0, 0)
    vsf_ctx.set_tuning()


def check_knn(ctx, Q, T):
    ei, ed = native.knn2_hamming(Q, T)
    idx, dist = ctx.knn2(Q, T)
    assert ctx.last_engine >= 2, "the tensor-core engine did not run"
    np.testing.assert_array_equal(idx, ei)
    np.testing.assert_array_equal(dist, ed)


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4]
                                        for p in glob.glob(os.path.join(GOLD, "knn_*_32.npz")) +
                                        glob.glob(os.path.join(GOLD, "knn_kat_*.npz"))))
def test_tc_knn2_against_opencv_golden(tc_ctx, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    if g["Q"].shape[1] != 32:
        pytest.skip("tensor-core engine: 32-byte descriptors")
    idx, dist = tc_ctx.knn2(g["Q"], g["T"])
    assert tc_ctx.last_engine >= 2
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(dist, g["dist"])


@pytest.mark.parametrize("split", [0, 1, 2, 3, 5, 8, 32])
def test_tc_knn2_c2_all_splits(tc_ctx, split):
    Q, T = synth.descriptor_pair(2000, 2000, seed=0)
    tc_ctx.set_tuning(-1, split, 0, -1)
    check_knn(tc_ctx, Q, T)


@pytest.mark.parametrize("split", [0, 1, 7])
def test_tc_knn2_ties_lowest_train_index(tc_ctx, split):
    # many equidistant neighbours, also inside one 16-row bucket and across buckets / tiles
    Q, T = synth.tie_pair(1500, 2300, seed=3)
    tc_ctx.set_tuning(-1, split, 0, -1)
    check_knn(tc_ctx, Q, T)
    idx, _ = tc_ctx.knn2(Q, T)
    assert (idx[:, 0] < idx[:, 1]).sum() > 100


def test_tc_knn2_duplicate_rows_everywhere(tc_ctx):
    # all train rows identical: neighbours must be rows 0 and 1 for every query
    T = np.tile(np.arange(32, dtype=np.uint8), (777, 1))
    Q = synth.descriptor_pair(300, 10, seed=4)[0]
    idx, dist = tc_ctx.knn2(Q, T)
    assert (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all() and (dist[:, 0] == dist[:, 1]).all()


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (3, 2), (31, 33), (33, 31), (257, 129), (1000, 17),
                                   (17, 1000), (129, 4097), (256, 256), (255, 257), (513, 15), (300, 16),
                                   (300, 31), (64, 511), (64, 513)])
def test_tc_knn2_ragged_shapes(tc_ctx, nq, nt):
    Q, T = synth.descriptor_pair(nq, nt, seed=nq * 7 + nt)
    check_knn(tc_ctx, Q, T)


def test_tc_knn2_extreme_distances(tc_ctx):
    # complements (distance 256), identical rows (distance 0) and all-zero / all-one rows
    rng = np.random.default_rng(11)
    T = rng.integers(0, 256, (600, 32), dtype=np.uint8)
    T[5] = 0
    T[77] = 255
    Q = np.concatenate([T[:100] ^ np.uint8(255), T[100:200], np.zeros((3, 32), np.uint8),
                        np.full((3, 32), 255, np.uint8)])
    check_knn(tc_ctx, Q, T)


def test_tc_knn2_empty_inputs(tc_ctx):
    Q, T = synth.descriptor_pair(40, 30, seed=1)
    idx, dist = tc_ctx.knn2(Q[:0], T)
    assert idx.shape == (0, 2)
    idx, dist = tc_ctx.knn2(Q, T[:0])
    assert (idx == -1).all() and (dist == -1).all()
    assert len(tc_ctx.get_matches(Q, T[:0], RATIO)) == 0
    assert len(tc_ctx.get_matches(Q, T[:1], RATIO)) == 0      # quirk Q6
    assert len(tc_ctx.get_matches(Q[:0], T, RATIO)) == 0


def test_tc_knn2_full_size_properties(tc_ctx):
    """C5's frame size (20000 rows) against the C oracle + size-independent properties."""
    Q, T = synth.descriptor_pair(20000, 20000, seed=12, planted=0.5)
    check_knn(tc_ctx, Q, T)
    si, sd = tc_ctx.knn2(T, T)
    np.testing.assert_array_equal(si[:, 0], np.arange(20000))
    assert (sd[:, 0] == 0).all() and (sd[:, 1] > 0).all()


@pytest.mark.parametrize("ratio", [RATIO, 0.8, 1.0, 0.0, 0.3333333333333333])
def test_tc_get_matches_against_oracle(tc_ctx, ratio):
    Q, T = synth.descriptor_pair(2000, 2000, seed=2, max_flips=60)
    got = tc_ctx.get_matches(Q, T, ratio)
    assert tc_ctx.last_engine >= 2
    np.testing.assert_array_equal(got, native.get_matches(Q, T, ratio))


def test_tc_engines_agree_with_popc_engine(vsf_ctx):
    Q, T = synth.descriptor_pair(3000, 5000, seed=21)
    res = []
    for e in (1, 2, 3):
        vsf_ctx.set_engine(e, 0)
        res.append(vsf_ctx.knn2(Q, T) + (vsf_ctx.get_matches(Q, T, RATIO),))
        assert vsf_ctx.last_engine == e
    vsf_ctx.set_engine(0, 0)
    for r in res[1:]:
        np.testing.assert_array_equal(r[0], res[0][0])
        np.testing.assert_array_equal(r[1], res[0][1])
        np.testing.assert_array_equal(r[2], res[0][2])


def test_tc_automatic_engine_choice(vsf_ctx):
    vsf_ctx.set_engine(0, 0)
    Q, T = synth.descriptor_pair(300, 300, seed=1)
    vsf_ctx.knn2(Q, T)
    assert vsf_ctx.last_engine == 1            # small batch: POPC kernel (launch-bound either way)
    Q, T = synth.descriptor_pair(5000, 5000, seed=1)
    vsf_ctx.knn2(Q, T)
    assert vsf_ctx.last_engine == 2            # large batch: tensor cores


@pytest.mark.parametrize("engine", ENGINES)
def test_tc_window_sequence_and_feature_matches(engine):
    import vision_slam_frontend_b200 as vsf
    n, W = 1500, 4
    frames = [synth.synth_pose(n - 37 * (p % 3), p, 150, 99) for p in range(8)]
    bp = restate.BEST_PERCENT
    with vsf.Context(device=0, max_features=4096, desc_bytes=32, window=W) as ctx:
        ctx.set_engine(engine, 0)
        live = []
        for p, D in enumerate(frames):
            got = ctx.window_match(D, RATIO)
            assert ctx.last_engine == engine or not live
            for (fid, m), (pid, past) in zip(got, live):
                assert fid == pid
                np.testing.assert_array_equal(m, native.get_matches(past, D, RATIO))
            for sort_mode in (0, 1):
                fm = ctx.window_feature_matches(D, RATIO, float(bp), sort_mode)
                for (fid, pairs), (pid, past) in zip(fm, live):
                    m = native.get_matches(past, D, RATIO)
                    keep = restate.num_good_matches(len(m), bp)
                    order = restate.sort_order_stdsort(m) if sort_mode == 1 else restate.sort_order_stable(m)
                    exp = m[order][:keep]
                    np.testing.assert_array_equal(pairs[:, 0], exp["queryIdx"].astype(np.uint64))
                    np.testing.assert_array_equal(pairs[:, 1], exp["trainIdx"].astype(np.uint64))
            ctx.window_commit(p, len(D))
            if len(live) >= W:
                live.pop(0)
            live.append((p, D))


@pytest.mark.parametrize("engine", ENGINES)
def test_tc_observe_features_sequence(engine):
    """The fused ObserveImage path with device-side row counts (compacted frames) on the tensor cores."""
    import vision_slam_frontend_b200 as vsf
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    W = 3
    frames = synth.stereo_sequence(5, 2000, seed=9)
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order="stable")
    with vsf.Context(device=0, max_features=4096, desc_bytes=32, window=W) as ctx:
        ctx.set_engine(engine, 0)
        for p, (kl, dl, kr, dr) in enumerate(frames):
            past = list(fo.frame_list)
            r = fo.observe_features(kl, dl, kr, dr)
            got = ctx.observe_features(p, kl, dl, kr, dr, F, P1, P2, RATIO)
            assert ctx.last_engine == engine
            sm = r.stereo_matches
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][r.stereo_keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][r.stereo_keep])
            for (fid, m), pf in zip(got["window"], past):
                assert fid == pf.frame_ID
                np.testing.assert_array_equal(m, native.get_matches(pf.descriptors, r.left.descriptors, RATIO))
            tm = native.get_matches(r.right.descriptors, r.left.descriptors, RATIO)
            np.testing.assert_array_equal(got["tri_matches"], tm)


def test_tc_window_c4_full_size_device_entry_point(tc_ctx):
    import torch
    n, W, stride, seed = 5000, 10, 500, 77
    buf = torch.empty((W + 1, n, 32), dtype=torch.uint8, device="cuda")
    tc_ctx.synth_sequence_device(buf.data_ptr(), n, 0, W + 1, stride, seed)
    base = buf.data_ptr()
    tc_ctx.window_match_device([base + j * n * 32 for j in range(W)], [n] * W, base + W * n * 32, n, RATIO)
    assert tc_ctx.last_engine >= 2
    got = tc_ctx.fetch_window(W)
    host = buf.cpu().numpy()
    total = 0
    for j in range(W):
        np.testing.assert_array_equal(got[j], native.get_matches(host[j], host[W], RATIO))
        total += len(got[j])
    assert total > 10000


@pytest.mark.parametrize("width", [61, 64])
def test_tc_wide_descriptors_against_opencv_golden(width):
    """64-byte rows (AKAZE's 61 bytes zero-padded, BRISK's 64) on the tensor cores
    (csrc/knn2_tc64_kernel.cu): the OpenCV golden vectors, bit-exact; the e4m3 engine does not
    exist for this width."""
    import vision_slam_frontend_b200 as vsf
    g = np.load(os.path.join(GOLD, f"knn_planted_{width}.npz"))
    with vsf.Context(device=0, max_features=4096, desc_bytes=width, window=2) as ctx:
        with pytest.raises(vsf.VsfError):
            ctx.set_engine(3, 0)
        ctx.set_engine(2, 0)
        idx, dist = ctx.knn2(g["Q"], g["T"])
        assert ctx.last_engine == 2
        np.testing.assert_array_equal(idx, g["idx"])
        np.testing.assert_array_equal(dist, g["dist"])
        ctx.set_engine(1, 0)
        i1, d1 = ctx.knn2(g["Q"], g["T"])
        np.testing.assert_array_equal(i1, g["idx"])
        np.testing.assert_array_equal(d1, g["dist"])


@pytest.mark.parametrize("width", [61, 64])
@pytest.mark.parametrize("split", [0, 1, 3, 8])
def test_tc_wide_descriptors_ragged_sizes(width, split):
    import vision_slam_frontend_b200 as vsf
    with vsf.Context(device=0, max_features=5200, desc_bytes=width, window=2) as ctx:
        ctx.set_engine(2, 0)
        ctx.set_tuning(-1, split, 0, -1)
        for (nq, nt, seed) in [(128, 128, 1), (129, 300, 3), (1000, 33, 5), (77, 1, 6), (300, 4097, 8), (2300, 2100, 4),
                               (5000, 5000, 7)]:
            Q, T = synth.descriptor_pair(nq, nt, width=width, seed=seed)
            ei, ed = native.knn2_hamming(Q, T)
            idx, dist = ctx.knn2(Q, T)
            assert ctx.last_engine == 2
            np.testing.assert_array_equal(idx, ei)
            np.testing.assert_array_equal(dist, ed)
            np.testing.assert_array_equal(ctx.get_matches(Q, T, RATIO), native.get_matches(Q, T, RATIO))
        Q, T = synth.tie_pair(900, 1100, width=width)      # adversarial ties: lowest train index twice
        ei, ed = native.knn2_hamming(Q, T)
        idx, dist = ctx.knn2(Q, T)
        np.testing.assert_array_equal(idx, ei)
        np.testing.assert_array_equal(dist, ed)


@pytest.mark.parametrize("sort_mode", [0, 1])
def test_tc_wide_descriptors_window_and_pipeline(sort_mode):
    """61-byte frames through the blocking window call and the pipelined frame stream on the
    tensor cores: lists as the reference's loop produces them (src/slam_frontend.cc:424-434)."""
    import vision_slam_frontend_b200 as vsf
    rng = np.random.default_rng(3)
    n, W, width = 1400, 3, 61
    base = rng.integers(0, 256, (n + 800, width), dtype=np.uint8)
    frames = []
    for p in range(7):
        D = base[100 * p: 100 * p + n - 37 * (p % 3)].copy()
        flips = rng.integers(0, 8 * width, (len(D), 6))
        for j in range(flips.shape[1]):
            D[np.arange(len(D)), flips[:, j] // 8] ^= (1 << (flips[:, j] % 8)).astype(np.uint8)
        frames.append(D[rng.permutation(len(D))])
    bp = restate.BEST_PERCENT
    with vsf.Context(device=0, max_features=2048, desc_bytes=width, window=W) as ctx:
        ctx.set_engine(2, 0)
        live, expected = [], []
        for p, D in enumerate(frames):
            if ctx.window_in_flight() == vsf.PIPELINE_DEPTH:
                break
            ctx.window_submit(p, D, RATIO, float(bp), sort_mode, False)
            expected.append((p, list(live), D))
            if len(live) >= W:
                live.pop(0)
            live.append((p, D))
        assert ctx.last_engine == 2
        while expected:
            fid, got = ctx.window_collect()
            efid, elive, D = expected.pop(0)
            assert fid == efid and len(got) == len(elive)
            for (pfid, pairs), (pid, past) in zip(got, elive):
                assert pfid == pid
                m = native.get_matches(past, D, RATIO)
                keep = restate.num_good_matches(len(m), bp)
                order = restate.sort_order_stdsort(m) if sort_mode == 1 else restate.sort_order_stable(m)
                exp = m[order][:keep]
                assert len(pairs) == keep and keep > 100
                np.testing.assert_array_equal(pairs[:, 0], exp["queryIdx"].astype(np.uint64))
                np.testing.assert_array_equal(pairs[:, 1], exp["trainIdx"].astype(np.uint64))


def test_tc_per_kernel_times(tc_ctx):
    Q, T = synth.descriptor_pair(5000, 5000, seed=3)
    tc_ctx.set_profile(True)
    try:
        tc_ctx.knn2(Q, T)
        ms = tc_ctx.last_kernel_times()
    finally:
        tc_ctx.set_profile(False)
    assert len(ms) == 4 and all(v >= 0 for v in ms) and ms[1] > 0


def test_tc_window_c5_size_long_pieces(tc_ctx):
    """BASELINE config 5's frame size (20000 features), 4 past frames: 25 000 tile slots, which
    puts the work partition in its "few long pieces per block" regime (>= 128 tiles per SM)."""
    import torch
    ctx = tc_ctx
    n, W, stride, seed = 20000, 4, 2000, 31
    buf = torch.empty((W + 1, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(buf.data_ptr(), n, 0, W + 1, stride, seed)
    base = buf.data_ptr()
    # ragged: the train frame has 19 877 rows, the query frames 20 000, 19 999, ...
    nt = n - 123
    nq = [n - j for j in range(W)]
    ctx.window_match_device([base + j * n * 32 for j in range(W)], nq, base + W * n * 32, nt, RATIO)
    assert ctx.last_engine >= 2
    got = ctx.fetch_window(W)
    host = buf.cpu().numpy()
    total = 0
    for j in range(W):
        np.testing.assert_array_equal(got[j], native.get_matches(host[j][:nq[j]], host[W][:nt], RATIO))
        total += len(got[j])
    assert total > 10000


def test_tc_wide_descriptors_observe_features_sequence():
    """The fused ObserveImage path on 61-byte descriptors with the tensor-core engine: device-side
    row counts (compacted frames), two train frames in one batch, stereo filter, triangulation."""
    import vision_slam_frontend_b200 as vsf
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    W = 3
    frames = synth.stereo_sequence(5, 2000, seed=9, width=61)
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order="stable")
    with vsf.Context(device=0, max_features=4096, desc_bytes=61, window=W) as ctx:
        ctx.set_engine(2, 0)
        for p, (kl, dl, kr, dr) in enumerate(frames):
            past = list(fo.frame_list)
            r = fo.observe_features(kl, dl, kr, dr)
            got = ctx.observe_features(p, kl, dl, kr, dr, F, P1, P2, RATIO)
            assert ctx.last_engine == 2
            sm = r.stereo_matches
            np.testing.assert_array_equal(got["kept_left"], sm["queryIdx"][r.stereo_keep])
            np.testing.assert_array_equal(got["kept_right"], sm["trainIdx"][r.stereo_keep])
            for (fid, m), pf in zip(got["window"], past):
                assert fid == pf.frame_ID
                np.testing.assert_array_equal(m, native.get_matches(pf.descriptors, r.left.descriptors, RATIO))
            tm = native.get_matches(r.right.descriptors, r.left.descriptors, RATIO)
            np.testing.assert_array_equal(got["tri_matches"], tm)


@pytest.mark.parametrize("width", [32, 61, 64])
def test_tc_random_shapes_against_popc_engine(width):
    """Random (queries, train rows) shapes, forced onto the tensor cores with random piece counts:
    indices, distances and ratio survivors equal to the POPC engine's (which the other tests pin
    to the oracle and to OpenCV)."""
    import vision_slam_frontend_b200 as vsf
    rng = np.random.default_rng(100 + width)
    with vsf.Context(device=0, max_features=3000, desc_bytes=width, window=2) as ctx:
        for it in range(24):
            nq = int(rng.integers(1, 3000))
            nt = int(rng.integers(1, 3000)) if it % 5 else int(rng.integers(1, 40))
            Q, T = synth.descriptor_pair(nq, nt, width=width, seed=int(rng.integers(1 << 30)))
            if it % 3 == 0:                       # duplicate rows: ties between distant train rows
                T[rng.integers(0, nt, nt // 2 + 1)] = T[rng.integers(0, nt, nt // 2 + 1)]
            ctx.set_engine(1, 0)
            ctx.set_tuning()
            ri, rd = ctx.knn2(Q, T)
            rm = ctx.get_matches(Q, T, RATIO)
            ctx.set_engine(2, 0)
            ctx.set_tuning(-1, int(rng.choice([0, 0, 1, 2, 3, 6, 11])), 0, -1)
            gi, gd = ctx.knn2(Q, T)
            gm = ctx.get_matches(Q, T, RATIO)
            assert ctx.last_engine == 2
            np.testing.assert_array_equal(gi, ri, err_msg=f"nq={nq} nt={nt}")
            np.testing.assert_array_equal(gd, rd, err_msg=f"nq={nq} nt={nt}")
            np.testing.assert_array_equal(gm, rm)
        ctx.set_tuning()
