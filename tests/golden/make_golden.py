"""Freeze golden vectors from the real OpenCV functions the reference calls.

Run once in the build container (cv2 4.13.0 importable):
    python tests/golden/make_golden.py
The reference (/root/reference) has no fixtures of its own for this path and its
C++ cannot be built here; its arithmetic lives in OpenCV
(src/slam_frontend.cc:525-527 knnMatch, :152-156 triangulatePoints, :334-339
undistortPoints).  These files pin the oracle restatements (oracle/restate.py,
oracle/oracle_knn.c) and, through them, the CUDA path, to OpenCV's own outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cv2  # noqa: E402

import synth  # noqa: E402
from oracle import cv2_ref  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, cv2_version=np.array(cv2.__version__), **arrays)
    print(f"{name}: {os.path.getsize(path)} bytes")


def knn_case(name, Q, T):
    idx, dist = cv2_ref.knn2_hamming(Q, T)
    save(name, Q=Q, T=T, idx=idx, dist=dist)


def main():
    # C2 shape, reduced: planted near-duplicates (ratio passes) + uniform rows
    knn_case("knn_planted_32", *synth.descriptor_pair(300, 280, 32, seed=0))
    # adversarial ties: lowest train index must win for both neighbours
    knn_case("knn_ties_32", *synth.tie_pair(200, 260, 32, seed=3))
    # AKAZE (61 bytes) and BRISK/FREAK (64 bytes) widths
    knn_case("knn_planted_61", *synth.descriptor_pair(150, 170, 61, seed=4))
    knn_case("knn_planted_64", *synth.descriptor_pair(150, 170, 64, seed=5))
    # ragged edges: fewer than two train rows
    Q, T = synth.descriptor_pair(9, 1, 32, seed=6)
    knn_case("knn_nt1_32", Q, T)
    # known-answer tests probed in SURVEY.md section 8(c)
    q = np.zeros((1, 32), np.uint8)
    t = np.zeros((5, 32), np.uint8)
    t[0, 0] = 1; t[2, 5] = 2; t[4, 31] = 128          # distances 1,0,1,0,1
    knn_case("knn_kat_alternating", q, t)
    knn_case("knn_kat_all_equal", np.zeros((3, 32), np.uint8), np.full((6, 32), 255, np.uint8))

    # C3 shape, reduced: KITTI-size rig, noisy correspondences
    P1, P2 = synth.kitti_projections()
    kl, dl, kr, dr, X, perm = synth.stereo_frame(240, seed=1, bad_geometry=0.0)
    ok = perm >= 0
    x1 = np.stack([kl["x"][ok], kl["y"][ok]], 1)
    x2 = np.stack([kr["x"][perm[ok]], kr["y"][perm[ok]]], 1)
    X4 = cv2_ref.triangulate_points(P1, P2, x1, x2)
    save("triangulate_kitti", P1=P1, P2=P2, x1=x1, x2=x2, X4=X4)
    # the reference's own PointGrey rig (src/slam_frontend.cc:565-611)
    Kl = np.array([[527.873518, 0, 482.823413], [0, 527.276819, 298.033945], [0, 0, 1]])
    Kr = np.array([[530.158021, 0, 475.540633], [0, 529.682234, 299.995465], [0, 0, 1]])
    A = np.array([[0.999593617649873, 0.021411909431148, -0.018818333830411, -0.131707087331978],
                  [-0.021140534893290, 0.999671312094879, 0.014503294761121, 0.003232397463343],
                  [0.019122691705565, -0.014099571235136, 0.999717722536176, -0.001146108483477]])
    PL = (Kl.astype(np.float32) @ np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32))
    PR = (Kr.astype(np.float32) @ A.astype(np.float32))
    rng = np.random.default_rng(11)
    Xw = np.stack([rng.uniform(-3, 3, 200), rng.uniform(-1, 1, 200), rng.uniform(0.5, 20, 200),
                   np.ones(200)])
    a = PL.astype(np.float64) @ Xw
    b = PR.astype(np.float64) @ Xw
    x1 = ((a[:2] / a[2]).T + rng.normal(0, 0.3, (200, 2))).astype(np.float32)
    x2 = ((b[:2] / b[2]).T + rng.normal(0, 0.3, (200, 2))).astype(np.float32)
    save("triangulate_pointgrey", P1=PL, P2=PR, x1=x1, x2=x2,
         X4=cv2_ref.triangulate_points(PL, PR, x1, x2))

    # N1: undistortPoints with the reference's left-camera coefficients (:565-573)
    dist = np.array([-0.153137, 0.075666, -0.000227, -0.000320, 0.0], np.float32)
    px = np.stack([rng.uniform(0, 960, 300), rng.uniform(0, 600, 300)], 1).astype(np.float32)
    save("undistort_pointgrey", K=Kl.astype(np.float32), dist=dist, px=px,
         out=cv2_ref.undistort_points(px, Kl, dist))


if __name__ == "__main__":
    main()
