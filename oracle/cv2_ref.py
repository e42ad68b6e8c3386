"""The real OpenCV functions the reference calls, through the Python wheel.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's hot arithmetic lives in OpenCV (un-vendored; pinned EXACT 3.2.0
at /root/reference/CMakeLists.txt:21).  The C++ library is absent from this
image, but `opencv-python-headless 4.13.0` exposes the same entry points:
  cv::BFMatcher::knnMatch      <- src/slam_frontend.cc:525-527
  cv::triangulatePoints        <- src/slam_frontend.cc:152-156
  cv::undistortPoints          <- src/slam_frontend.cc:334-339
These wrappers are used (1) by tests/golden/make_golden.py to freeze golden
vectors, (2) by tests to cross-check the restatement live when cv2 imports,
(3) by bench.py as the multi-threaded CPU baseline ("the reference's OpenCV CPU
path timed on the box's own host cores", BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np

from .restate import DMATCH_DTYPE, ratio_pass


def available() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def set_threads(n: int) -> int:
    import cv2
    cv2.setNumThreads(int(n))
    return cv2.getNumThreads()


def version() -> str:
    import cv2
    return cv2.__version__


def knn2_hamming(Q: np.ndarray, T: np.ndarray):
    """cv2.BFMatcher(NORM_HAMMING).knnMatch(Q, T, k=2) -> (idx, dist) int32
    (nq,2), -1 where the neighbour does not exist."""
    import cv2
    nq = len(Q)
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq == 0 or len(T) == 0:
        return idx, dist
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(
        np.ascontiguousarray(Q), np.ascontiguousarray(T), 2)
    for i, row in enumerate(m):
        for k, dm in enumerate(row[:2]):
            assert dm.queryIdx == i and dm.imgIdx == 0
            assert float(dm.distance) == int(dm.distance)
            idx[i, k] = dm.trainIdx
            dist[i, k] = int(dm.distance)
    return idx, dist


def knn_match_raw(Q: np.ndarray, T: np.ndarray):
    """The bare call, for timing."""
    import cv2
    return cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(Q, T, 2)


def get_matches(Q: np.ndarray, T: np.ndarray, nn_match_ratio: float) -> np.ndarray:
    """Frontend::GetMatches (src/slam_frontend.cc:521-538) on top of real cv2."""
    idx, dist = knn2_hamming(Q, T)
    if len(T) < 2 or len(Q) == 0:
        return np.zeros(0, dtype=DMATCH_DTYPE)
    keep = ratio_pass(dist[:, 0], dist[:, 1], nn_match_ratio)
    out = np.zeros(int(keep.sum()), dtype=DMATCH_DTYPE)
    out["queryIdx"] = np.nonzero(keep)[0]
    out["trainIdx"] = idx[keep, 0]
    out["distance"] = dist[keep, 0].astype(np.float32)
    return out


def triangulate_points(P1, P2, x1, x2) -> np.ndarray:
    """cv2.triangulatePoints with float32 inputs; x1,x2 are (n,2)."""
    import cv2
    x1 = np.ascontiguousarray(np.asarray(x1, np.float32).reshape(-1, 2).T)
    x2 = np.ascontiguousarray(np.asarray(x2, np.float32).reshape(-1, 2).T)
    if x1.shape[1] == 0:
        return np.zeros((4, 0), np.float32)
    return cv2.triangulatePoints(np.asarray(P1, np.float32).reshape(3, 4),
                                 np.asarray(P2, np.float32).reshape(3, 4), x1, x2)


def undistort_points(px, K, dist) -> np.ndarray:
    import cv2
    px = np.asarray(px, np.float32).reshape(-1, 1, 2)
    K = np.asarray(K, np.float32).reshape(3, 3)
    out = cv2.undistortPoints(px, K, np.asarray(dist, np.float32).reshape(-1, 1),
                              None, K)
    return out.reshape(-1, 2)
