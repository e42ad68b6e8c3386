"""The C++ slam::Frontend mirror (csrc/frontend/) against the FrontendOracle state
machine: same call sequence as the reference's ObserveImage, entered after feature
extraction.  Runs on the GPU through libvsf_frontend.so -> libvsf_cuda.so."""
import numpy as np
import pytest

import synth
from oracle import native, restate

pytestmark = pytest.mark.gpu

RATIO = restate.NN_MATCH_RATIO


@pytest.mark.parametrize("exact", [True, False])
def test_frontend_sequence_matches_oracle(exact):
    from vision_slam_frontend_b200.frontend import Frontend, parse_slam_problem
    P1, P2 = synth.kitti_projections()
    F = synth.kitti_fundamental()
    K = synth.KITTI_K.astype(np.float32)
    dist = np.array([-0.153137, 0.075666, -0.000227, -0.000320, 0.0], np.float32)
    W = 3
    frames = synth.stereo_sequence(6, 1500, seed=21)
    order = "stdsort" if exact else "stable"
    fo = restate.FrontendOracle(P1, P2, F, frame_life=W, order=order)
    results = []
    with Frontend(max_features=2048, desc_bytes=32, frame_life=W, P_left=P1, P_right=P2, fundamental=F,
                  K_left=K, dist_left=dist, exact_std_sort=exact) as fe:
        # the first frame is gated out: prev odometry == current odometry (OdomCheck)
        fe.observe_odometry([0, 0, 0], [1, 0, 0, 0], 9.0)
        assert fe.observe_features(*frames[0]) is False and fe.num_poses == 0
        added = []
        for p, (kl, dl, kr, dr) in enumerate(frames):
            fe.observe_odometry([0.5 * (p + 1), 0, 0], [1, 0, 0, 0], 10.0 + p)
            ok = fe.observe_features(kl, dl, kr, dr, 10.0 + p)
            added.append(ok)
            results.append(fo.observe_features(kl, dl, kr, dr))
            assert fe.stereo_threshold.view(np.uint32) == np.float32(fo.stereo_ambig_constraint).view(np.uint32)
        assert all(added) and fe.num_poses == len(frames)

        # vision factors: same order, same pose ids, identical FeatureMatch lists
        got = fe.vision_factors()
        exp = fo.vision_factors
        assert len(got) == len(exp) == sum(min(p, W) for p in range(len(frames)))
        for (a, b, pairs), e in zip(got, exp):
            assert (a, b) == (e.pose_idx_initial, e.pose_idx_current)
            np.testing.assert_array_equal(pairs, e.feature_matches)
        assert sum(len(p) for _, _, p in got) > 1000

        # nodes: features = (i, undistorted pixel, points[i]) with the reference's indexing
        for p, r in enumerate(results):
            node = fe.node(p)
            assert node["timestamp"] == 10.0 + p and len(node["pixel"]) == len(r.features_pixel)
            np.testing.assert_allclose(node["pixel"], restate.undistort_points(r.features_pixel, K, dist), atol=2e-3)
            e3 = r.features_point3d
            np.testing.assert_array_equal(np.isnan(node["point3d"]), np.isnan(e3))
            okm = ~np.isnan(e3).any(1)
            rel = np.abs(node["point3d"][okm] - e3[okm]).max(1) / np.abs(e3[okm]).max(1)
            assert rel.max() < 1e-4, rel.max()
            np.testing.assert_allclose(node["loc"], [0.5 * (p + 1), 0, 0], atol=1e-6)

        # odometry factors between consecutive poses
        of = fe.odometry_factors()
        assert [(i, j) for i, j, _, _ in of] == [(p - 1, p) for p in range(1, len(frames))]
        for _, _, t, q in of:
            np.testing.assert_allclose(t, [0.5, 0, 0], atol=1e-6)
            np.testing.assert_allclose(q, [0, 0, 0, 1], atol=1e-6)

        # ROS1 wire format round trip
        msg = parse_slam_problem(fe.serialize_problem())
        assert len(msg["nodes"]) == len(frames) and len(msg["vision_factors"]) == len(got)
        for (a, b, pairs), (a2, b2, pairs2) in zip(got, msg["vision_factors"]):
            assert (a, b) == (a2, b2)
            np.testing.assert_array_equal(pairs, pairs2)
        n3 = msg["nodes"][3]
        node = fe.node(3)
        assert n3["id"] == 3 and len(n3["features"]) == len(node["pixel"])
        fid, pix, p3 = n3["features"][5]
        assert fid == 5 and pix[2] == 0.0
        np.testing.assert_array_equal(np.float32(pix[:2]), node["pixel"][5])
        assert len(msg["odometry_factors"]) == len(frames) - 1


def test_frontend_get_matches_operator_boundary():
    from vision_slam_frontend_b200.frontend import Frontend
    Q, T = synth.descriptor_pair(900, 800, seed=5)
    with Frontend(max_features=1024, desc_bytes=32) as fe:
        np.testing.assert_array_equal(fe.get_matches(Q, T, RATIO), native.get_matches(Q, T, RATIO))


def test_frontend_default_config_is_the_reference_rig():
    from vision_slam_frontend_b200.frontend import default_config
    cfg = default_config()
    np.testing.assert_allclose(cfg["P_left"][0], [527.873518, 0, 482.823413, 0], rtol=1e-6)
    np.testing.assert_allclose(cfg["K_left"][1], [0, 527.276819, 298.033945], rtol=1e-6)
    np.testing.assert_allclose(cfg["dist_left"], [-0.153137, 0.075666, -0.000227, -0.000320, 0], rtol=1e-5)
    # textbook epipolar geometry: x_l^T F x_r = 0 for a projected 3-D point
    Kl = cfg["K_left"].astype(np.float64)
    A = np.array([[0.999593617649873, 0.021411909431148, -0.018818333830411, -0.131707087331978],
                  [-0.021140534893290, 0.999671312094879, 0.014503294761121, 0.003232397463343],
                  [0.019122691705565, -0.014099571235136, 0.999717722536176, -0.001146108483477]])
    X = np.array([0.7, -0.2, 5.0])
    xl = cfg["P_left"].astype(np.float64) @ np.append(X, 1)
    xr = cfg["P_right"].astype(np.float64) @ np.append(X, 1)
    xl, xr = xl / xl[2], xr / xr[2]
    Fn = cfg["fundamental"].astype(np.float64)
    assert abs(xl @ Fn @ xr) / np.abs(Fn).max() < 1e-2
