"""Seeded synthetic inputs of the BASELINE.json shapes.

Shared by the tests, the golden-vector generator and bench.py's CPU legs.
Pure numpy; nothing here touches the product library or the oracle.
"""
from __future__ import annotations

import numpy as np

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


def flip_bits(rows: np.ndarray, max_flips: int, rng) -> np.ndarray:
    """Flip k ~ U{0..max_flips} random bit positions (with replacement) per row."""
    rows = rows.copy()
    n, w = rows.shape
    k = rng.integers(0, max_flips + 1, n)
    for i in range(n):
        for b in rng.integers(0, w * 8, k[i]):
            rows[i, b // 8] ^= np.uint8(1 << (b % 8))
    return rows


def descriptor_pair(nq: int, nt: int, width: int = 32, planted: float = 0.5,
                    max_flips: int = 20, seed: int = 0):
    """C2-style pair: T uniform; `planted` of the Q rows are a random T row with up
    to max_flips flipped bits, the rest uniform (uniform rows never pass the 0.6
    ratio test, planted ones almost always do)."""
    rng = np.random.default_rng(seed)
    T = rng.integers(0, 256, (nt, width), dtype=np.uint8)
    Q = rng.integers(0, 256, (nq, width), dtype=np.uint8)
    if nt > 0 and nq > 0:
        npl = int(nq * planted)
        rows = rng.permutation(nq)[:npl]
        Q[rows] = flip_bits(T[rng.integers(0, nt, npl)], max_flips, rng)
    return Q, T


def tie_pair(nq: int, nt: int, width: int = 32, distinct: int = 7, seed: int = 3):
    """Adversarial ties: train rows drawn from only `distinct` codes, so every query
    has many equidistant neighbours; the lowest train index must win twice."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 256, (distinct, width), dtype=np.uint8)
    T = codes[rng.integers(0, distinct, nt)]
    Q = flip_bits(codes[rng.integers(0, distinct, nq)], 3, rng)
    return Q, T


def make_keypoints(xy: np.ndarray) -> np.ndarray:
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    kp = np.zeros(len(xy), KEYPOINT_DTYPE)
    kp["x"], kp["y"] = xy[:, 0], xy[:, 1]
    kp["size"], kp["angle"], kp["class_id"] = 31.0, -1.0, -1
    return kp


# KITTI-odometry-style rectified stereo rig for 1241x376 images (C3)
KITTI_K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1]], np.float64)
KITTI_T = np.array([-0.5371657, 0.0, 0.0])      # X_right = X_left + t  (baseline 0.537 m)


def kitti_projections():
    P_left = (KITTI_K @ np.hstack([np.eye(3), np.zeros((3, 1))])).astype(np.float32)
    P_right = (KITTI_K @ np.hstack([np.eye(3), KITTI_T.reshape(3, 1)])).astype(np.float32)
    return P_left, P_right


def kitti_fundamental() -> np.ndarray:
    """F with x_left^T F x_right = 0 (the reference's convention), float32."""
    t = KITTI_T
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    Ki = np.linalg.inv(KITTI_K)
    F_std = Ki.T @ tx @ Ki              # x_r^T F_std x_l = 0
    F = F_std.T
    return (F / np.abs(F).max()).astype(np.float32)


def stereo_frame(n: int, seed: int = 1, outliers: float = 0.1, noise_px: float = 0.5,
                 width: int = 32, max_flips: int = 20, landmarks=None, bad_geometry: float = 0.05):
    """C3-style stereo pair: n 3-D points in front of the rig projected into both
    cameras with Gaussian pixel noise; right descriptors = left with up to
    max_flips flipped bits, right order shuffled; `outliers` of the right
    features are unrelated (random descriptor, random pixel) and `bad_geometry`
    of them keep their descriptor but sit 30 px off the epipolar line (they pass
    the ratio test and must be removed by the epipolar filter).
    Returns kp_left, desc_left, kp_right, desc_right, X (n,3) ground truth,
    perm (right index of left feature i, -1 for outliers)."""
    rng = np.random.default_rng(seed)
    P_left, P_right = kitti_projections()
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-3, 3, n), rng.uniform(4, 80, n)], 1)
    Xh = np.hstack([X, np.ones((n, 1))])
    pl = (P_left.astype(np.float64) @ Xh.T).T
    pr = (P_right.astype(np.float64) @ Xh.T).T
    xl = (pl[:, :2] / pl[:, 2:3] + rng.normal(0, noise_px, (n, 2))).astype(np.float32)
    xr = (pr[:, :2] / pr[:, 2:3] + rng.normal(0, noise_px, (n, 2))).astype(np.float32)
    if landmarks is None:
        dl = rng.integers(0, 256, (n, width), dtype=np.uint8)
    else:
        dl = flip_bits(landmarks, max_flips // 2, rng)
    dr = flip_bits(dl, max_flips, rng)
    nout = int(n * outliers)
    out_rows = rng.permutation(n)[:nout]
    dr[out_rows] = rng.integers(0, 256, (nout, width), dtype=np.uint8)
    xr[out_rows] = np.stack([rng.uniform(0, 1241, nout), rng.uniform(0, 376, nout)], 1)
    rng2 = np.random.default_rng(seed + 12345)
    bad_rows = rng2.permutation(n)[:int(n * bad_geometry)]
    xr[bad_rows, 1] += (30.0 * rng2.choice([-1.0, 1.0], len(bad_rows))).astype(np.float32)
    order = rng.permutation(n)           # right feature j is left feature order[j]
    dr, xr = dr[order], xr[order]
    perm = np.empty(n, np.int64)
    perm[order] = np.arange(n)
    perm[out_rows] = -1
    return make_keypoints(xl), dl, make_keypoints(xr), dr, X.astype(np.float32), perm


def stereo_sequence(n_poses: int, n: int, seed: int = 5, overlap: float = 0.8, width: int = 32):
    """A short sequence of stereo frames whose descriptors overlap from pose to
    pose (pose p observes landmarks [s*p, s*p+n), s = (1-overlap)*n)."""
    rng = np.random.default_rng(seed)
    s = max(1, int(round((1 - overlap) * n)))
    pool = rng.integers(0, 256, (s * n_poses + n, width), dtype=np.uint8)
    frames = []
    for p in range(n_poses):
        lm = pool[s * p: s * p + n][rng.permutation(n)]
        frames.append(stereo_frame(n, seed=seed * 1000 + p, landmarks=lm, width=width)[:4])
    return frames


# ---- numpy twin of the CUDA sequence generator (csrc/aux_kernels.cu) ----------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def mix64(x):
    with np.errstate(over="ignore"):
        x = (np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def synth_pose(n: int, pose: int, stride: int, seed: int, width: int = 32) -> np.ndarray:
    """Descriptors (n, width) uint8 of one pose of the synthetic sequence; identical, bit for
    bit, to vsf_synth_sequence_device on a context of that descriptor width (32-byte device
    rows for width <= 32, 64-byte rows above; bytes beyond `width` are zero on the device and
    not returned here)."""
    import math
    words = 8 if width <= 32 else 16
    seed = np.uint64(seed)
    with np.errstate(over="ignore"):
        hp = int(mix64(seed ^ (np.uint64(0xA5A5A5A5) + np.uint64(pose) * np.uint64(0x100000001B3))))
        a = (hp % n) | 1
        while math.gcd(a, n) != 1:
            a += 2
        b = (hp >> 32) % n
        i = np.arange(n, dtype=np.uint64)
        perm = (np.uint64(a) * i + np.uint64(b)) % np.uint64(n)
        L = np.uint64(stride) * np.uint64(pose) + perm
        w = np.arange(words, dtype=np.uint64)
        code = (mix64(seed ^ (L[:, None] * np.uint64(words) + w[None, :])) & np.uint64(0xFFFFFFFF))
        c = (np.uint64(pose) * np.uint64(n) + i)[:, None] * np.uint64(words) + w[None, :]
        ns = ~seed
        r0 = mix64(ns ^ (c * np.uint64(3)))
        r1 = mix64(ns ^ (c * np.uint64(3) + np.uint64(1)))
        r2 = mix64(ns ^ (c * np.uint64(3) + np.uint64(2)))
        lo = np.uint64(0xFFFFFFFF)
        flip = (r0 & lo) & (r0 >> np.uint64(32)) & (r1 & lo) & (r1 >> np.uint64(32)) & (r2 & lo)
        out = (code ^ flip).astype(np.uint32)
    return np.ascontiguousarray(out.view(np.uint8).reshape(n, 4 * words)[:, :width])
