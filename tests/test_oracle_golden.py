"""The oracle restatements against the frozen OpenCV outputs (tests/golden/) and,
when cv2 imports, against live OpenCV on fresh seeded inputs.  CPU only."""
import glob
import os

import numpy as np
import pytest

import synth
from oracle import cv2_ref, native, restate

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNN_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "knn_*.npz")))


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name", KNN_CASES)
def test_numpy_knn_matches_opencv_golden(name):
    g = load(name)
    idx, dist = restate.knn2_hamming(g["Q"], g["T"])
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(dist, g["dist"])


@pytest.mark.parametrize("name", KNN_CASES)
def test_c_knn_matches_opencv_golden(name):
    g = load(name)
    idx, dist = native.knn2_hamming(g["Q"], g["T"])
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_array_equal(dist, g["dist"])


def test_known_answers_from_survey():
    # q = 0, train distances {1,0,1,0,1} -> neighbours (idx 1, d 0), (idx 3, d 0)
    g = load("knn_kat_alternating")
    assert g["idx"].tolist() == [[1, 3]] and g["dist"].tolist() == [[0, 0]]
    # all equal -> (0, d), (1, d): lowest train index wins both slots
    g = load("knn_kat_all_equal")
    assert g["idx"].tolist() == [[0, 1]] * 3 and g["dist"].tolist() == [[256, 256]] * 3
    # fewer than two train rows -> the second neighbour does not exist
    g = load("knn_nt1_32")
    assert (g["idx"][:, 1] == -1).all() and (g["idx"][:, 0] == 0).all()


def test_edge_shapes():
    Q, T = synth.descriptor_pair(5, 4, seed=9)
    for fn in (restate.knn2_hamming, native.knn2_hamming):
        i, d = fn(Q[:0], T)
        assert i.shape == (0, 2)
        i, d = fn(Q, T[:0])
        assert (i == -1).all() and (d == -1).all()
    assert len(restate.get_matches(Q, T[:1])) == 0      # quirk Q6: nt < 2 -> nothing passes
    assert len(native.get_matches(Q, T[:1], 0.6)) == 0


def test_ratio_constant_and_integer_equivalence():
    # 0.6f widened to double is 10066330 / 2^24; the double compare equals the
    # exact integer compare for every reachable pair of distances
    assert restate.NN_MATCH_RATIO == 10066330 / 2 ** 24
    d1, d2 = np.meshgrid(np.arange(513), np.arange(513), indexing="ij")
    a = restate.ratio_pass(d1.ravel(), d2.ravel(), restate.NN_MATCH_RATIO)
    b = (d1.ravel().astype(np.int64) << 24) < 10066330 * d2.ravel().astype(np.int64)
    np.testing.assert_array_equal(a, b)


def test_num_good_matches_is_floor():
    for n in list(range(0, 300)) + [999, 5000, 19999, 20000]:
        assert restate.num_good_matches(n, np.float32(0.3)) == int(np.floor(np.float32(n) * np.float32(0.3)))
        assert restate.num_good_matches(n, np.float32(1.0)) == n


def test_get_matches_three_ways():
    Q, T = synth.descriptor_pair(400, 380, seed=21)
    a = restate.get_matches(Q, T, restate.NN_MATCH_RATIO)
    b = native.get_matches(Q, T, restate.NN_MATCH_RATIO)
    np.testing.assert_array_equal(a, b)
    assert len(a) > 100 and (np.diff(a["queryIdx"]) > 0).all() and (a["imgIdx"] == 0).all()


@pytest.mark.parametrize("name", ["triangulate_kitti", "triangulate_pointgrey"])
def test_triangulation_matches_opencv_golden(name):
    g = load(name)
    X4 = restate.triangulate_points(g["P1"], g["P2"], g["x1"], g["x2"])
    assert X4.dtype == np.float32 and X4.shape == g["X4"].shape
    a, b = restate.dehomogenize(X4), restate.dehomogenize(g["X4"])
    rel = np.abs(a - b).max(1) / np.abs(b).max(1)
    assert rel.max() < 1e-5, rel.max()          # north_star tolerance is 1e-4
    # homogeneous columns agree up to sign and are unit norm
    sgn = np.sign((X4 * g["X4"]).sum(0))
    np.testing.assert_allclose(X4 * sgn, g["X4"], atol=2e-6)


def test_undistort_matches_opencv_golden():
    g = load("undistort_pointgrey")
    out = restate.undistort_points(g["px"], g["K"], g["dist"])
    np.testing.assert_allclose(out, g["out"], atol=2e-3)


def test_stdsort_order_is_a_sort_and_differs_only_in_ties():
    Q, T = synth.descriptor_pair(600, 500, seed=33, max_flips=6)
    m = restate.get_matches(Q, T)
    o_std, o_stable = restate.sort_order_stdsort(m), restate.sort_order_stable(m)
    assert sorted(o_std.tolist()) == list(range(len(m)))
    np.testing.assert_array_equal(m["distance"][o_std], m["distance"][o_stable])
    assert (np.diff(m["distance"][o_std]) >= 0).all()


def test_epipolar_residual_zero_on_true_correspondences():
    P1, P2 = synth.kitti_projections()
    kl, dl, kr, dr, X, perm = synth.stereo_frame(100, seed=2, noise_px=0.0, outliers=0.0, bad_geometry=0.0)
    F = synth.kitti_fundamental()
    xl = np.stack([kl["x"], kl["y"]], 1)
    xr = np.stack([kr["x"][perm], kr["y"][perm]], 1)
    c = restate.epipolar_residual(xl, xr, F)
    bad = restate.epipolar_residual(xl, xr[::-1].copy(), F)
    assert c.max() < 1e-3 and np.median(bad) > 10 * c.max()


def test_frontend_oracle_window_and_threshold():
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    fo = restate.FrontendOracle(P1, P2, F, frame_life=3)
    frames = synth.stereo_sequence(5, 120, seed=7)
    seen = []
    for p, (kl, dl, kr, dr) in enumerate(frames):
        r = fo.observe_features(kl, dl, kr, dr)
        assert len(r.vision_factors) == min(p, 3)
        assert [v.pose_idx_initial for v in r.vision_factors] == list(range(max(0, p - 3), p))
        assert all(v.pose_idx_current == p for v in r.vision_factors)
        assert len(r.left.keypoints) == len(r.right.keypoints) == int(r.stereo_keep.sum())
        assert len(r.points) == len(r.tri_matches) <= len(r.left.keypoints)
        seen.append(fo.stereo_ambig_constraint)
        if p > 0:
            assert sum(len(v.feature_matches) for v in r.vision_factors) > 0
    assert seen[0] != restate.STEREO_AMBIG_INIT and np.isfinite(seen).all()
    assert len(fo.frame_list) == 3 and [f.frame_ID for f in fo.frame_list] == [2, 3, 4]


def test_synth_twin_is_deterministic_and_overlapping():
    a = synth.synth_pose(500, 3, 50, 42)
    b = synth.synth_pose(500, 3, 50, 42)
    np.testing.assert_array_equal(a, b)
    nxt = synth.synth_pose(500, 4, 50, 42)
    m = restate.get_matches(a, nxt)
    assert 0.75 * 450 < len(m) <= 450          # 90 % of the landmarks are shared


@pytest.mark.skipif(not cv2_ref.available(), reason="cv2 not importable")
def test_live_opencv_cross_check():
    Q, T = synth.descriptor_pair(700, 650, seed=77)
    i0, d0 = cv2_ref.knn2_hamming(Q, T)
    i1, d1 = restate.knn2_hamming(Q, T)
    i2, d2 = native.knn2_hamming(Q, T)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)
    np.testing.assert_array_equal(i0, i2)
    np.testing.assert_array_equal(d0, d2)
    Q, T = synth.tie_pair(300, 400, seed=78)
    np.testing.assert_array_equal(cv2_ref.knn2_hamming(Q, T)[0], native.knn2_hamming(Q, T)[0])
