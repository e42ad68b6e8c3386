"""numpy restatement of the reference's matching / stereo / triangulation path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it follows; paths are relative to /root/reference.

The arithmetic of `cv::BFMatcher::knnMatch` and `cv::triangulatePoints` lives
in OpenCV (un-vendored, pinned EXACT 3.2.0 at CMakeLists.txt:21).  Their
published algorithms are restated here and pinned against the outputs of the
importable OpenCV 4.13 wheel (tests/golden/, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------------------
# Plain-data mirrors of the OpenCV / reference types that cross the boundary.
# ---------------------------------------------------------------------------

# cv::KeyPoint memory layout (28 bytes): Point2f pt; float size, angle,
# response; int octave, class_id.  Frame::keypoints_ (src/slam_frontend.h:107).
KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

# cv::DMatch memory layout (16 bytes): int queryIdx, trainIdx, imgIdx; float
# distance.  What Frontend::GetMatches returns (src/slam_frontend.cc:521-538).
DMATCH_DTYPE = np.dtype(
    [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"),
     ("distance", "<f4")])

# FrontendConfig defaults (src/slam_frontend.cc:550-559).
NN_MATCH_RATIO = float(np.float32(0.6))   # float member widened to double (:555, :523)
BEST_PERCENT = np.float32(0.3)            # :554
FRAME_LIFE = 10                           # :556
STEREO_AMBIG_INIT = np.float32(10000.0)   # file-scope static (:353)
STEREO_PADDING = np.float32(2.0)          # :392


def make_keypoints(xy: np.ndarray) -> np.ndarray:
    """Keypoint array whose .pt is `xy` (n,2); other fields like cv::KeyPoint()."""
    xy = np.asarray(xy, dtype=np.float32).reshape(-1, 2)
    kp = np.zeros(len(xy), dtype=KEYPOINT_DTYPE)
    kp["x"], kp["y"] = xy[:, 0], xy[:, 1]
    kp["size"] = 31.0
    kp["angle"] = -1.0
    kp["class_id"] = -1
    return kp


# ---------------------------------------------------------------------------
# a1: cv::BFMatcher(NORM_HAMMING)::knnMatch(query, train, matches, k=2)
#     call site src/slam_frontend.cc:525-527, matcher built at :247.
# ---------------------------------------------------------------------------

def _as_u64_rows(D: np.ndarray) -> np.ndarray:
    """View an (n, bytes) uint8 descriptor matrix as (n, ceil(bytes/8)) uint64,
    zero-padding the tail (zero padding does not change Hamming distances;
    AKAZE descriptors are 61 bytes)."""
    D = np.ascontiguousarray(D, dtype=np.uint8)
    n, w = D.shape
    w8 = (w + 7) // 8 * 8
    if w8 != w:
        P = np.zeros((n, w8), dtype=np.uint8)
        P[:, :w] = D
        D = P
    return D.view(np.uint64)


def hamming_matrix(Q: np.ndarray, T: np.ndarray) -> np.ndarray:
    """All-pairs Hamming distances, int32 (nq, nt)."""
    q, t = _as_u64_rows(Q), _as_u64_rows(T)
    assert q.shape[1] == t.shape[1], "descriptor widths differ"
    return np.bitwise_count(q[:, None, :] ^ t[None, :, :]).sum(-1).astype(np.int32)


def knn2_hamming(Q: np.ndarray, T: np.ndarray, chunk: int = 256
                 ) -> Tuple[np.ndarray, np.ndarray]:
    """k=2 brute-force Hamming nearest neighbours.

    OpenCV's BFMatcher (batchDistance with K=2, no mask, no cross-check) scans
    the train rows in increasing index order and inserts with strict `<`, so
    among equal distances the LOWEST train index wins, for the first and for
    the second neighbour; i.e. the result is the lexicographic top-2 of
    (distance, trainIdx).  np.argmin returns the first minimum, which is the
    same rule.  Returns (idx, dist), both int32 (nq, 2); entries that do not
    exist (nt < 2) are -1 (OpenCV returns shorter inner vectors there).
    """
    nq, nt = len(Q), len(T)
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq == 0 or nt == 0:
        return idx, dist
    for s in range(0, nq, chunk):
        d = hamming_matrix(Q[s:s + chunk], T)
        rows = np.arange(len(d))
        i0 = d.argmin(1)
        idx[s:s + chunk, 0] = i0
        dist[s:s + chunk, 0] = d[rows, i0]
        if nt >= 2:
            d[rows, i0] = np.iinfo(np.int32).max
            i1 = d.argmin(1)
            idx[s:s + chunk, 1] = i1
            dist[s:s + chunk, 1] = d[rows, i1]
    return idx, dist


# ---------------------------------------------------------------------------
# a2: Frontend::GetMatches  (src/slam_frontend.cc:521-538)
# ---------------------------------------------------------------------------

def ratio_pass(d1: np.ndarray, d2: np.ndarray, nn_match_ratio: float) -> np.ndarray:
    """`dist1 < nn_match_ratio * dist2` (:533): float distances widened to
    double, `nn_match_ratio` is a double parameter (:523)."""
    return d1.astype(np.float64) < np.float64(nn_match_ratio) * d2.astype(np.float64)


def get_matches(Q: np.ndarray, T: np.ndarray,
                nn_match_ratio: float = NN_MATCH_RATIO) -> np.ndarray:
    """Frontend::GetMatches: kNN(k=2) then Lowe ratio; survivors in ascending
    queryIdx order, imgIdx = 0.  The reference reads matches[i][1]
    unconditionally (:532), undefined for nt < 2 (quirk Q6); here and in the
    CUDA path nt < 2 means "no match passes"."""
    idx, dist = knn2_hamming(Q, T)
    if len(T) < 2 or len(Q) == 0:
        return np.zeros(0, dtype=DMATCH_DTYPE)
    keep = ratio_pass(dist[:, 0], dist[:, 1], nn_match_ratio)
    out = np.zeros(int(keep.sum()), dtype=DMATCH_DTYPE)
    out["queryIdx"] = np.nonzero(keep)[0]
    out["trainIdx"] = idx[keep, 0]
    out["distance"] = dist[keep, 0].astype(np.float32)
    return out


# ---------------------------------------------------------------------------
# a3: Frontend::GetFeatureMatches  (src/slam_frontend.cc:282-309)
# ---------------------------------------------------------------------------

def num_good_matches(n: int, best_percent) -> int:
    """`const int num_good_matches = matches.size() * config_.best_percent_`
    (:290): size_t -> float, float multiply, truncation to int."""
    return int(np.float32(n) * np.float32(best_percent))


def sort_order_stable(matches: np.ndarray) -> np.ndarray:
    """Documented-deviation order: ascending (distance, position); what the
    device-side sort produces.  Differs from std::sort only inside groups of
    equal distance."""
    return np.argsort(matches["distance"], kind="stable")


_STDSORT = None


def _stdsort_lib():
    global _STDSORT
    if _STDSORT is None:
        from . import build as _b
        _STDSORT = ctypes.CDLL(_b.ensure_built()["stdsort"])
        _STDSORT.oracle_stdsort_order.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _STDSORT.oracle_stdsort_order.restype = None
    return _STDSORT


def sort_order_stdsort(matches: np.ndarray) -> np.ndarray:
    """Permutation applied by `std::sort(matches.begin(), matches.end())`
    (:289) with cv::DMatch::operator< (distance only).  std::sort is not
    stable; the permutation is whatever libstdc++'s introsort does, so it is
    obtained by running libstdc++ itself (oracle/stdsort_oracle.cc)."""
    n = len(matches)
    order = np.empty(n, dtype=np.int32)
    if n:
        d = np.ascontiguousarray(matches["distance"], dtype=np.float32)
        _stdsort_lib().oracle_stdsort_order(d.ctypes.data, n, order.ctypes.data)
    return order


@dataclasses.dataclass
class Frame:
    """slam::Frame (src/slam_frontend.h:100-114; ctor src/slam_frontend.cc:511-519)."""
    keypoints: np.ndarray            # KEYPOINT_DTYPE (n,)
    descriptors: np.ndarray          # uint8 (n, bytes)
    frame_ID: int
    is_initial: np.ndarray = None    # bool (n,)
    initial_ids: np.ndarray = None   # int64 (n,)

    def __post_init__(self):
        n = len(self.keypoints)
        self.descriptors = np.ascontiguousarray(self.descriptors, dtype=np.uint8)
        if self.descriptors.ndim != 2:
            self.descriptors = self.descriptors.reshape(n, -1)
        if self.is_initial is None:
            self.is_initial = np.ones(n, dtype=bool)          # :517
        if self.initial_ids is None:
            self.initial_ids = np.full(n, -1, dtype=np.int64)  # :518


@dataclasses.dataclass
class VisionFactor:
    """slam_types::VisionFactor (src/slam_types.h:91-108)."""
    pose_idx_initial: int
    pose_idx_current: int
    feature_matches: np.ndarray      # uint64 (m, 2): initial, current


def get_feature_matches(past: Frame, curr: Frame,
                        nn_match_ratio: float = NN_MATCH_RATIO,
                        best_percent=BEST_PERCENT,
                        order: str = "stdsort",
                        return_matches: bool = False):
    """Frontend::GetFeatureMatches: GetMatches(past=query, curr=train) ->
    sort by distance -> keep the first int(n*best_percent) -> FeatureMatch
    (queryIdx -> initial, trainIdx -> current) + is_initial_ book-keeping."""
    m = get_matches(past.descriptors, curr.descriptors, nn_match_ratio)   # :287-288
    perm = sort_order_stdsort(m) if order == "stdsort" else sort_order_stable(m)
    m = m[perm]                                                            # :289
    m = m[:num_good_matches(len(m), best_percent)]                         # :290-291
    pairs = np.stack([m["queryIdx"].astype(np.uint64),
                      m["trainIdx"].astype(np.uint64)], axis=1).reshape(-1, 2)
    for q, t in zip(m["queryIdx"], m["trainIdx"]):                         # :293-306
        if curr.is_initial[t]:
            curr.is_initial[t] = False
            curr.initial_ids[t] = (past.frame_ID if past.is_initial[q]
                                   else past.initial_ids[q])
    vf = VisionFactor(past.frame_ID, curr.frame_ID, pairs)                 # :308
    return (vf, m) if return_matches else vf


# ---------------------------------------------------------------------------
# a5: Frontend::RemoveAmbigStereo  (src/slam_frontend.cc:353-398)
# ---------------------------------------------------------------------------

def epipolar_residual(xl: np.ndarray, xr: np.ndarray, F: np.ndarray, order: int = 0) -> np.ndarray:
    """`(left_ph.transpose() * config_.fundamental * right_ph).norm()` (:380-381)
    in float32 without FMA contraction: v = l^T F (three 3-term dot products),
    c = v . r (one more), then the norm of the 1x1 result = sqrt(c*c).

    The summation order inside a 3-term dot product is Eigen's, and Eigen is
    not vendored by the reference (nor installed here), so it cannot be pinned:
      order 0: (c0 + c1) + c2  - Eigen 3.2's unrolled coefficient product
               (product_coeff_impl<DefaultTraversal, 2>: res(0..1) + c2);
      order 1: c0 + (c1 + c2)  - Eigen 3.3's `.sum()` via redux_novec_unroller,
               which splits a range of 3 into 1 + 2.
    The two differ by at most one ulp of c; the device implements both
    (VSF_OPT_RESIDUAL_ORDER) and is bit-identical to this function for each."""
    f = np.asarray(F, dtype=np.float32).reshape(3, 3)
    xl = np.asarray(xl, dtype=np.float32).reshape(-1, 2)
    xr = np.asarray(xr, dtype=np.float32).reshape(-1, 2)
    one = np.float32(1.0)
    l = [xl[:, 0], xl[:, 1], np.full(len(xl), one, np.float32)]
    r = [xr[:, 0], xr[:, 1], np.full(len(xr), one, np.float32)]

    def dot3(a0, a1, a2):
        if order == 0:
            return ((a0 + a1).astype(np.float32) + a2).astype(np.float32)
        return (a0 + (a1 + a2).astype(np.float32)).astype(np.float32)

    with np.errstate(over="ignore", invalid="ignore"):
        v = [dot3((l[0] * f[0, j]).astype(np.float32), (l[1] * f[1, j]).astype(np.float32),
                  (l[2] * f[2, j]).astype(np.float32)) for j in range(3)]
        c = dot3((v[0] * r[0]).astype(np.float32), (v[1] * r[1]).astype(np.float32),
                 (v[2] * r[2]).astype(np.float32))
        return np.sqrt((c * c).astype(np.float32)).astype(np.float32)


def remove_ambig_stereo(left: Frame, right: Frame, stereo_matches: np.ndarray,
                        F: np.ndarray, thresh: np.float32, residual_order: int = 0,
                        hold_on_empty: bool = False):
    """Returns (left', right', new_thresh, residuals, keep_mask).

    Survivors keep match order (ascending left index); both frames are rebuilt
    index-aligned with fresh is_initial_/initial_ids_ (:396-397).  The new
    threshold is mean(residual over ALL matches) + 2 (:392-394), accumulated in
    float32 in match order; with zero matches it is 0/0 = NaN exactly as in
    the reference (quirk Q4)."""
    qi, ti = stereo_matches["queryIdx"], stereo_matches["trainIdx"]
    xl = np.stack([left.keypoints["x"][qi], left.keypoints["y"][qi]], 1)
    xr = np.stack([right.keypoints["x"][ti], right.keypoints["y"][ti]], 1)
    c = epipolar_residual(xl, xr, F, residual_order)
    avg = np.float32(0.0)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for v in c:                                   # :382 sequential float sum
            avg = np.float32(avg + v)
        keep = c <= np.float32(thresh)                # :383 (false for NaN)
        new_thresh = np.float32(np.float32(avg / np.float32(len(c))) + STEREO_PADDING)
    if hold_on_empty and len(c) == 0:                 # opt-in deviation (VSF_OPT_HOLD_THRESHOLD_ON_EMPTY)
        new_thresh = np.float32(thresh)
    L = Frame(left.keypoints[qi[keep]], left.descriptors[qi[keep]], left.frame_ID)
    R = Frame(right.keypoints[ti[keep]], right.descriptors[ti[keep]], right.frame_ID)
    return L, R, new_thresh, c, keep


def default_fundamental(K_left, K_right, R, t) -> np.ndarray:
    """The reference's own construction (src/slam_frontend.cc:635-644) reads
    A[3] of a 3-vector and is therefore undefined (quirk Q3).  The boundary
    takes F as an input; this helper builds the textbook matrix in the
    reference's convention `x_left^T F x_right = 0`, where X_right = R X_left + t:
    F = (K_r^-T [t]x R K_l^-1)^T, computed in double, rounded to float32."""
    K_left, K_right = np.asarray(K_left, np.float64), np.asarray(K_right, np.float64)
    R, t = np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    F_std = np.linalg.inv(K_right).T @ tx @ R @ np.linalg.inv(K_left)
    return F_std.T.astype(np.float32)


# ---------------------------------------------------------------------------
# a6: cv::triangulatePoints + Frontend::Calculate3DPoints (:117-173)
# ---------------------------------------------------------------------------

def triangulate_points(P1, P2, x1, x2) -> np.ndarray:
    """cv::triangulatePoints(P1, P2, pts1, pts2) for float32 inputs.

    Published algorithm (OpenCV calib3d triangulate.cpp): per point build the
    4x4 matrix with two rows per view, `x*P[2,:] - P[0,:]` and
    `y*P[2,:] - P[1,:]`, in double; SVD; the homogeneous point is the right
    singular vector of the smallest singular value; result stored as float32
    4xN.  The sign of the singular vector is arbitrary (it cancels in xyz/w).
    x1, x2: (n,2) float32 pixel coordinates.
    """
    P1 = np.asarray(P1, dtype=np.float32).astype(np.float64).reshape(3, 4)
    P2 = np.asarray(P2, dtype=np.float32).astype(np.float64).reshape(3, 4)
    x1 = np.asarray(x1, dtype=np.float32).astype(np.float64).reshape(-1, 2)
    x2 = np.asarray(x2, dtype=np.float32).astype(np.float64).reshape(-1, 2)
    n = len(x1)
    if n == 0:
        return np.zeros((4, 0), dtype=np.float32)
    A = np.empty((n, 4, 4))
    A[:, 0] = x1[:, 0:1] * P1[2] - P1[0]
    A[:, 1] = x1[:, 1:2] * P1[2] - P1[1]
    A[:, 2] = x2[:, 0:1] * P2[2] - P2[0]
    A[:, 3] = x2[:, 1:2] * P2[2] - P2[1]
    _, _, vt = np.linalg.svd(A)
    return vt[:, 3, :].T.astype(np.float32)


def dehomogenize(X4: np.ndarray) -> np.ndarray:
    """`Vector3f(x, y, z) / w` in float32 (:161-164). Returns (n,3)."""
    X4 = np.asarray(X4, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (X4[:3] / X4[3]).T.astype(np.float32)


def calculate_3d_points(left: Frame, right: Frame, P_left, P_right,
                        nn_match_ratio: float = NN_MATCH_RATIO,
                        order: str = "stdsort"):
    """Frontend::Calculate3DPoints: GetFeatureMatches(right=query/initial,
    left=train/current) with best_percent forced to 1.0 (:129-132), gather
    pixel pairs in sorted-match order (:136-150), triangulate (:152-156),
    divide by w (:159-165).  Returns (points (m,3) f32, matches DMATCH (m,),
    X4 (4,m) f32).  Mutates left.is_initial / left.initial_ids exactly like
    the reference does (left is the 'current' frame of that call)."""
    vf, m = get_feature_matches(right, left, nn_match_ratio, np.float32(1.0),
                                order=order, return_matches=True)
    if len(m) == 0:                                   # :133-135
        return np.zeros((0, 3), np.float32), m, np.zeros((4, 0), np.float32)
    lt, rq = m["trainIdx"], m["queryIdx"]
    xl = np.stack([left.keypoints["x"][lt], left.keypoints["y"][lt]], 1)
    xr = np.stack([right.keypoints["x"][rq], right.keypoints["y"][rq]], 1)
    X4 = triangulate_points(P_left, P_right, xl, xr)
    return dehomogenize(X4), m, X4


# ---------------------------------------------------------------------------
# N1: cv::undistortPoints as called by Frontend::UndistortFeaturePoints
#     (src/slam_frontend.cc:323-351): R = empty, P = K_left.
# ---------------------------------------------------------------------------

def undistort_points(px: np.ndarray, K, dist, iters: int = 5) -> np.ndarray:
    """Published algorithm of cv::undistortPoints (imgproc undistort.cpp) for
    the 5-coefficient model (k1,k2,p1,p2,k3): normalise with K, run the fixed
    point iteration x <- (x0 - tangential(x)) / radial(x) `iters` times in
    double (OpenCV's default criteria: 5 iterations), re-project with P = K,
    store float32.  px: (n,2) float32."""
    K = np.asarray(K, dtype=np.float32).astype(np.float64).reshape(3, 3)
    k = np.zeros(5)
    d = np.asarray(dist, dtype=np.float32).astype(np.float64).ravel()
    k[:len(d)] = d[:5]
    k1, k2, p1, p2, k3 = k
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    px = np.asarray(px, dtype=np.float32).astype(np.float64).reshape(-1, 2)
    x0 = (px[:, 0] - cx) / fx
    y0 = (px[:, 1] - cy) / fy
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        # OpenCV clamps a negative inverse-distortion factor by restoring x0,y0.
        bad = icdist < 0
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
        x = np.where(bad, x0, xn)
        y = np.where(bad, y0, yn)
    out = np.stack([x * fx + cx, y * fy + cy], 1)
    return out.astype(np.float32)


# ---------------------------------------------------------------------------
# a4 + ObserveImage ordering (src/slam_frontend.cc:400-472), from the point
# where features have been extracted (extraction is the input producer and is
# out of scope, BASELINE.json north_star).
# ---------------------------------------------------------------------------

@dataclasses.dataclass
class ObserveResult:
    stereo_matches: np.ndarray        # DMATCH, L->R after ratio (query order)
    stereo_residuals: np.ndarray      # f32 per stereo match
    stereo_keep: np.ndarray           # bool per stereo match
    left: Frame                       # compacted current frame
    right: Frame                      # compacted right frame
    vision_factors: List[VisionFactor]
    tri_matches: np.ndarray           # DMATCH, R->L, sorted order
    points: np.ndarray                # (m',3) f32 in sorted R->L match order
    features_pixel: np.ndarray        # (M,2) f32 (distorted pixels, pre-N1)
    features_point3d: np.ndarray      # (M,3) f32, reference indexing (quirk Q5)


class FrontendOracle:
    """State machine equal to slam::Frontend for the matching path: sliding
    window `frame_list_` (:193, :467-470), `curr_frame_ID_`, the file-scope
    adaptive stereo threshold (:353), accumulated vision factors (:432)."""

    def __init__(self, P_left, P_right, fundamental,
                 nn_match_ratio: float = NN_MATCH_RATIO,
                 best_percent=BEST_PERCENT, frame_life: int = FRAME_LIFE,
                 order: str = "stdsort", residual_order: int = 0, hold_on_empty: bool = False):
        self.residual_order = int(residual_order)
        self.hold_on_empty = bool(hold_on_empty)
        self.P_left = np.asarray(P_left, np.float32).reshape(3, 4)
        self.P_right = np.asarray(P_right, np.float32).reshape(3, 4)
        self.F = np.asarray(fundamental, np.float32).reshape(3, 3)
        self.nn_match_ratio = float(nn_match_ratio)
        self.best_percent = np.float32(best_percent)
        self.frame_life = int(frame_life)
        self.order = order
        self.stereo_ambig_constraint = STEREO_AMBIG_INIT
        self.curr_frame_ID = 0
        self.frame_list: List[Frame] = []
        self.vision_factors: List[VisionFactor] = []

    def observe_features(self, kp_left, desc_left, kp_right, desc_right) -> ObserveResult:
        curr = Frame(kp_left, desc_left, self.curr_frame_ID)             # :411
        right = Frame(kp_right, desc_right, self.curr_frame_ID)          # :412
        stereo = get_matches(curr.descriptors, right.descriptors,
                             self.nn_match_ratio)                        # :414-416
        curr, right, new_t, resid, keep = remove_ambig_stereo(
            curr, right, stereo, self.F, self.stereo_ambig_constraint,
            self.residual_order, self.hold_on_empty)                     # :417
        self.stereo_ambig_constraint = new_t
        factors = []
        for past in self.frame_list:                                     # :424-434
            vf = get_feature_matches(past, curr, self.nn_match_ratio,
                                     self.best_percent, order=self.order)
            factors.append(vf)
            self.vision_factors.append(vf)
        points, tri_m, _ = calculate_3d_points(
            curr, right, self.P_left, self.P_right, self.nn_match_ratio,
            order=self.order)                                            # :437
        M = len(curr.keypoints)
        pix = np.stack([curr.keypoints["x"], curr.keypoints["y"]], 1).astype(np.float32)
        # :438-442 indexes points[i] by keypoint index i although `points` is
        # in sorted-match order and may be shorter (quirk Q5, out-of-bounds
        # read in the reference).  Defined behaviour here: NaN where i >= len.
        p3 = np.full((M, 3), np.nan, dtype=np.float32)
        k = min(M, len(points))
        p3[:k] = points[:k]
        self.curr_frame_ID += 1                                          # :457
        if len(self.frame_list) >= self.frame_life:                      # :467-469
            self.frame_list.pop(0)
        self.frame_list.append(curr)                                     # :470
        return ObserveResult(stereo, resid, keep, curr, right, factors,
                             tri_m, points, pix, p3)
