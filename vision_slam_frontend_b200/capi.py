"""ctypes view of the C ABI in include/vsf.h (one method per entry point).

No fallback of any kind: `load_library()` raises if `libvsf_cuda.so` has not
been built, and `Context(...)` raises if `vsf_create` cannot get a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# VSF_LIB_PATH: development knob to load an alternative build of the same library (kernel variants)
_LIB_PATH = os.environ.get("VSF_LIB_PATH") or os.path.join(_PKG, "libvsf_cuda.so")
_LIB = None

# cv::DMatch / cv::KeyPoint / slam_types::FeatureMatch layouts (include/vsf.h)
OPT_RESIDUAL_ORDER = 1            # VSF_OPT_RESIDUAL_ORDER
OPT_HOLD_THRESHOLD_ON_EMPTY = 2
OPT_POSE_GROUP = 3   # VSF_OPT_HOLD_THRESHOLD_ON_EMPTY
OPT_DEBUG_SORT_DEPTH = 100        # VSF_OPT_DEBUG_SORT_DEPTH
PIPELINE_DEPTH = 16   # VSF_PIPELINE_DEPTH (include/vsf.h): frames vsf_window_submit keeps in flight

DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"),
                         ("imgIdx", "<i4"), ("distance", "<f4")])
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"),
                           ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
FEATURE_MATCH_DTYPE = np.dtype([("feature_idx_initial", "<u8"),
                                ("feature_idx_current", "<u8")])

ERR_NAMES = {0: "VSF_OK", 1: "VSF_ERR_BAD_ARG", 2: "VSF_ERR_CAPACITY",
             3: "VSF_ERR_CUDA", 4: "VSF_ERR_STATE"}


class VsfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class _ObserveOut(C.Structure):
    _fields_ = [("kept_left", C.c_void_p), ("kept_right", C.c_void_p),
                ("n_kept", C.c_int), ("stereo_threshold_next", C.c_float),
                ("n_frames", C.c_int), ("frame_ids", C.c_void_p),
                ("window_counts", C.c_void_p), ("window_matches", C.c_void_p),
                ("n_tri", C.c_int), ("tri_matches", C.c_void_p),
                ("tri_X4", C.c_void_p), ("cap", C.c_int), ("xy_undist", C.c_void_p)]


class _ObserveParams(C.Structure):
    _fields_ = [("fundamental", C.c_void_p), ("P_left", C.c_void_p), ("P_right", C.c_void_p),
                ("K_left", C.c_void_p), ("dist_left", C.c_void_p), ("nn_match_ratio", C.c_double)]


def library_path() -> str:
    return _LIB_PATH


def load_library():
    """dlopen libvsf_cuda.so; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise ImportError(
            f"{_LIB_PATH} not built: run `python -m vision_slam_frontend_b200.build` "
            "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i, sz, d, f, u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_double, C.c_float, C.c_uint64
    sigs = {
        "vsf_create": ([i, i, i, i, C.POINTER(vp)], i),
        "vsf_destroy": ([vp], None),
        "vsf_last_error": ([vp], C.c_char_p),
        "vsf_version": ([], C.c_char_p),
        "vsf_set_stream": ([vp, vp], i),
        "vsf_synchronize": ([vp], i),
        "vsf_set_tuning": ([vp, i, i, i, i], i),
        "vsf_set_engine": ([vp, i, i], i),
        "vsf_last_engine": ([vp], i),
        "vsf_set_profile": ([vp, i], i),
        "vsf_last_kernel_times": ([vp, C.POINTER(f)], i),
        "vsf_knn2": ([vp, vp, i, sz, vp, i, sz, vp, vp], i),
        "vsf_get_matches": ([vp, vp, i, sz, vp, i, sz, d, vp, i, C.POINTER(i)], i),
        "vsf_window_push": ([vp, u64, vp, i, sz], i),
        "vsf_window_commit": ([vp, u64, i], i),
        "vsf_window_clear": ([vp], i),
        "vsf_window_size": ([vp], i),
        "vsf_window_match": ([vp, vp, i, sz, d, vp, vp, vp, i, C.POINTER(i)], i),
        "vsf_window_feature_matches": ([vp, vp, i, sz, d, f, i, vp, vp, vp, i, C.POINTER(i)], i),
        "vsf_window_submit": ([vp, u64, vp, i, sz, d, f, i, i], i),
        "vsf_window_collect": ([vp, C.POINTER(u64), vp, vp, vp, i, C.POINTER(i)], i),
        "vsf_window_in_flight": ([vp], i),
        "vsf_set_host_threads": ([vp, i], i),
        "vsf_window_last_transfer": ([vp, C.POINTER(sz), C.POINTER(sz)], i),
        "vsf_stereo_filter": ([vp, vp, vp, i, sz, vp, vp, i, sz, vp, d, vp, vp,
                               C.POINTER(i), vp, vp, C.POINTER(i)], i),
        "vsf_set_stereo_threshold": ([vp, f], i),
        "vsf_get_stereo_threshold": ([vp, C.POINTER(f)], i),
        "vsf_set_option": ([vp, i, i], i),
        "vsf_get_option": ([vp, i, C.POINTER(i)], i),
        "vsf_triangulate": ([vp, vp, vp, vp, vp, i, vp], i),
        "vsf_undistort_points": ([vp, vp, vp, vp, i, vp], i),
        "vsf_observe_features": ([vp, u64, vp, vp, i, sz, vp, vp, i, sz, vp, vp, vp, d,
                                  C.POINTER(_ObserveOut)], i),
        "vsf_device_row_bytes": ([vp], i),
        "vsf_window_match_device": ([vp, vp, vp, i, vp, i, d], i),
        "vsf_fetch_window": ([vp, i, vp, vp, i], i),
        "vsf_window_match_block_device": ([vp, vp, i, i, C.c_longlong, i, d], i),
        "vsf_window_run_sequence": ([vp, vp, i, i, C.c_longlong, i, d, f, i, i, vp, vp, i, i,
                                     C.POINTER(sz), C.POINTER(sz)], i),
        "vsf_synth_sequence_device": ([vp, vp, i, i, i, i, u64], i),
        "vsf_device_match_lists": ([vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i), C.POINTER(i)], i),
        "vsf_stream": ([vp], vp),
        "vsf_observe_submit": ([vp, u64, vp, vp, i, sz, vp, vp, i, sz, C.POINTER(_ObserveParams)], i),
        "vsf_observe_collect": ([vp, C.POINTER(u64), C.POINTER(_ObserveOut)], i),
        "vsf_observe_in_flight": ([vp], i),
        "vsf_probe_pipe": ([vp, i, i, C.POINTER(d)], i),
        "vsf_device_sm_count": ([vp], i),
        "vsf_debug_launch_count": ([vp], C.c_longlong),
        "vsf_debug_tc_trace": ([vp, vp, i, C.POINTER(i)], i),
        "vsf_debug_kernel_trace": ([vp, vp, i, C.POINTER(i)], i),
        "vsf_debug_tc_plan": ([i, i, i, i, C.c_longlong, C.c_longlong, vp], i),
        "vsf_debug_sort_prefix": ([vp, i, i], i),
        "vsf_debug_sort_prefix_depth": ([vp, i, i, i], i),
        "vsf_debug_sort_device": ([vp, vp, i, f, i, vp, C.POINTER(i)], i),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(L, name)   # AttributeError if the library lacks a declared symbol
        fn.argtypes = args
        fn.restype = res
    _LIB = L
    return L


EXPORTED_SYMBOLS = [
    "vsf_create", "vsf_destroy", "vsf_last_error", "vsf_version", "vsf_set_stream",
    "vsf_synchronize", "vsf_set_tuning", "vsf_set_engine", "vsf_last_engine", "vsf_set_profile",
    "vsf_last_kernel_times", "vsf_knn2", "vsf_get_matches", "vsf_window_push",
    "vsf_window_commit", "vsf_window_clear", "vsf_window_size", "vsf_window_match",
    "vsf_window_feature_matches", "vsf_window_submit", "vsf_window_collect",
    "vsf_window_in_flight", "vsf_set_host_threads", "vsf_window_last_transfer", "vsf_stereo_filter", "vsf_set_stereo_threshold",
    "vsf_get_stereo_threshold", "vsf_set_option", "vsf_get_option", "vsf_triangulate", "vsf_undistort_points", "vsf_observe_features",
    "vsf_device_row_bytes", "vsf_window_match_device", "vsf_fetch_window",
    "vsf_window_match_block_device", "vsf_window_run_sequence",
    "vsf_synth_sequence_device", "vsf_device_match_lists", "vsf_stream", "vsf_observe_submit",
    "vsf_observe_collect", "vsf_observe_in_flight", "vsf_probe_pipe", "vsf_device_sm_count", "vsf_debug_launch_count",
    "vsf_debug_tc_trace", "vsf_debug_kernel_trace", "vsf_debug_tc_plan", "vsf_debug_sort_prefix",
    "vsf_debug_sort_prefix_depth", "vsf_debug_sort_device",
]


def _u8rows(a: np.ndarray, width: Optional[int] = None) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("descriptors must be a 2-D uint8 array")
    if width is not None and a.shape[1] != width:
        raise ValueError("descriptor rows are %d bytes wide, the context was created for %d"
                         % (a.shape[1], width))
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class Context:
    """One vsf_ctx: one CUDA device, one stream, device-resident sliding window."""

    def __init__(self, device: int = 0, max_features: int = 8192, desc_bytes: int = 32,
                 window: int = 10):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.vsf_create(device, max_features, desc_bytes, window, C.byref(h))
        if rc != 0:
            raise VsfError(rc, "vsf_create failed (a CUDA device is required; there is no CPU path)")
        self._h = h
        self.device, self.max_features, self.desc_bytes, self.window = (
            device, max_features, desc_bytes, window)

    # -- plumbing ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.vsf_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise VsfError(rc, self._L.vsf_last_error(self._h).decode())

    @property
    def sm_count(self) -> int:
        return self._L.vsf_device_sm_count(self._h)

    @property
    def row_bytes(self) -> int:
        return self._L.vsf_device_row_bytes(self._h)

    def set_stream(self, cuda_stream: int):
        self._check(self._L.vsf_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self._check(self._L.vsf_synchronize(self._h))

    def set_tuning(self, popc_mode=-1, train_split=0, queries_per_thread=0, variant=-1):
        self._check(self._L.vsf_set_tuning(self._h, popc_mode, train_split, queries_per_thread,
                                           variant))

    def launch_count(self) -> int:
        """Kernels launched so far by the kNN paths of this ctx."""
        return int(self._L.vsf_debug_launch_count(self._h))

    def set_engine(self, engine: int = 0, flags: int = 0):
        """0 auto, 1 POPC pipe, 2 tensor cores (int8), 3 tensor cores (e4m3)."""
        self._check(self._L.vsf_set_engine(self._h, engine, flags))

    @property
    def last_engine(self) -> int:
        return self._L.vsf_last_engine(self._h)

    def set_profile(self, enabled: bool):
        self._check(self._L.vsf_set_profile(self._h, int(bool(enabled))))

    def last_kernel_times(self):
        """[expand, distance/selection kernel, refine, compaction] durations in ms."""
        ms = (C.c_float * 4)()
        self._check(self._L.vsf_last_kernel_times(self._h, ms))
        return [float(v) for v in ms]

    # -- a1 / a2 -------------------------------------------------------------------
    def knn2(self, Q: np.ndarray, T: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        Q, T = _u8rows(Q, self.desc_bytes), _u8rows(T, self.desc_bytes)
        idx = np.full((len(Q), 2), -1, np.int32)
        dist = np.full((len(Q), 2), -1, np.int32)
        self._check(self._L.vsf_knn2(self._h, _ptr(Q), len(Q), Q.strides[0], _ptr(T), len(T),
                                     T.strides[0], _ptr(idx), _ptr(dist)))
        return idx, dist

    def get_matches(self, Q: np.ndarray, T: np.ndarray, ratio: float) -> np.ndarray:
        Q, T = _u8rows(Q, self.desc_bytes), _u8rows(T, self.desc_bytes)
        out = np.zeros(max(len(Q), 1), DMATCH_DTYPE)
        n = C.c_int(0)
        self._check(self._L.vsf_get_matches(self._h, _ptr(Q), len(Q), Q.strides[0], _ptr(T),
                                            len(T), T.strides[0], float(ratio), _ptr(out),
                                            len(out), C.byref(n)))
        return out[:n.value].copy()

    # -- a4 ---------------------------------------------------------------------------
    def window_push(self, frame_id: int, D: np.ndarray):
        D = _u8rows(D, self.desc_bytes)
        self._check(self._L.vsf_window_push(self._h, frame_id, _ptr(D), len(D), D.strides[0]))

    def window_commit(self, frame_id: int, n: int):
        self._check(self._L.vsf_window_commit(self._h, frame_id, n))

    def window_clear(self):
        self._check(self._L.vsf_window_clear(self._h))

    def window_size(self) -> int:
        return self._L.vsf_window_size(self._h)

    def window_match(self, D: np.ndarray, ratio: float):
        """-> list of (frame_id, DMATCH array) per resident past frame, oldest first."""
        D = _u8rows(D, self.desc_bytes)
        cap = self.max_features
        fids = np.zeros(self.window, np.uint64)
        counts = np.zeros(self.window, np.int32)
        out = np.zeros((self.window, cap), DMATCH_DTYPE)
        nf = C.c_int(0)
        self._check(self._L.vsf_window_match(self._h, _ptr(D), len(D), D.strides[0], float(ratio),
                                             _ptr(fids), _ptr(counts), _ptr(out), cap, C.byref(nf)))
        return [(int(fids[j]), out[j, :counts[j]].copy()) for j in range(nf.value)]

    def window_feature_matches(self, D: np.ndarray, ratio: float, best_percent: float,
                               sort_mode: int = 1):
        """-> list of (frame_id, (m,2) uint64 [initial, current]) per past frame."""
        D = _u8rows(D, self.desc_bytes)
        cap = self.max_features
        fids = np.zeros(self.window, np.uint64)
        counts = np.zeros(self.window, np.int32)
        out = np.zeros((self.window, cap), FEATURE_MATCH_DTYPE)
        nf = C.c_int(0)
        self._check(self._L.vsf_window_feature_matches(
            self._h, _ptr(D), len(D), D.strides[0], float(ratio), float(best_percent), sort_mode,
            _ptr(fids), _ptr(counts), _ptr(out), cap, C.byref(nf)))
        res = []
        for j in range(nf.value):
            fm = out[j, :counts[j]]
            res.append((int(fids[j]), np.stack([fm["feature_idx_initial"],
                                                fm["feature_idx_current"]], 1).reshape(-1, 2)))
        return res

    # pipelined form: submit frames ahead, collect their FeatureMatch lists in order
    def window_submit(self, frame_id: int, D: np.ndarray, ratio: float, best_percent: float,
                      sort_mode: int = 1, pinned: bool = False):
        """pinned=True: D is page-locked, row-contiguous at the device row width and stays
        untouched until the frame is collected (VSF_SUBMIT_PINNED_DESC)."""
        D = _u8rows(D, self.desc_bytes)
        self._check(self._L.vsf_window_submit(self._h, frame_id, _ptr(D), len(D), D.strides[0],
                                              float(ratio), float(best_percent), sort_mode,
                                              1 if pinned else 0))

    def window_last_transfer(self):
        """(h2d_bytes, d2h_bytes) of the most recent window call."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        self._check(self._L.vsf_window_last_transfer(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_host_threads(self, n: int):
        self._check(self._L.vsf_set_host_threads(self._h, n))

    def window_in_flight(self) -> int:
        return self._L.vsf_window_in_flight(self._h)

    def window_collect(self):
        """-> (frame_id, [(past_frame_id, (m,2) uint64 [initial, current]), ...]) of the
        oldest submitted frame."""
        cap = self.max_features
        fids = np.zeros(self.window, np.uint64)
        counts = np.zeros(self.window, np.int32)
        out = np.zeros((self.window, cap), FEATURE_MATCH_DTYPE)
        nf = C.c_int(0)
        fid = C.c_uint64(0)
        self._check(self._L.vsf_window_collect(self._h, C.byref(fid), _ptr(fids), _ptr(counts),
                                               _ptr(out), cap, C.byref(nf)))
        res = []
        for j in range(nf.value):
            fm = out[j, :counts[j]]
            res.append((int(fids[j]), np.stack([fm["feature_idx_initial"],
                                                fm["feature_idx_current"]], 1).reshape(-1, 2)))
        return int(fid.value), res

    # -- a5 ---------------------------------------------------------------------------
    def stereo_filter(self, kp_left, desc_left, kp_right, desc_right, F, ratio: float):
        kl = np.ascontiguousarray(kp_left, KEYPOINT_DTYPE)
        kr = np.ascontiguousarray(kp_right, KEYPOINT_DTYPE)
        dl, dr = _u8rows(desc_left, self.desc_bytes), _u8rows(desc_right, self.desc_bytes)
        F = np.ascontiguousarray(F, np.float32).reshape(9)
        cap = max(len(kl), 1)
        kept_l = np.zeros(cap, np.int32)
        kept_r = np.zeros(cap, np.int32)
        sm = np.zeros(cap, DMATCH_DTYPE)
        resid = np.zeros(cap, np.float32)
        nk, ns = C.c_int(0), C.c_int(0)
        self._check(self._L.vsf_stereo_filter(
            self._h, _ptr(kl), _ptr(dl), len(kl), dl.strides[0], _ptr(kr), _ptr(dr), len(kr),
            dr.strides[0], _ptr(F), float(ratio), _ptr(kept_l), _ptr(kept_r), C.byref(nk),
            _ptr(sm), _ptr(resid), C.byref(ns)))
        return dict(kept_left=kept_l[:nk.value].copy(), kept_right=kept_r[:nk.value].copy(),
                    stereo_matches=sm[:ns.value].copy(), residuals=resid[:ns.value].copy())

    def set_stereo_threshold(self, v: float):
        self._check(self._L.vsf_set_stereo_threshold(self._h, float(v)))

    def set_option(self, option: int, value: int):
        self._check(self._L.vsf_set_option(self._h, option, value))

    def get_option(self, option: int) -> int:
        v = C.c_int(0)
        self._check(self._L.vsf_get_option(self._h, option, C.byref(v)))
        return v.value

    def get_stereo_threshold(self) -> np.float32:
        v = C.c_float(0)
        self._check(self._L.vsf_get_stereo_threshold(self._h, C.byref(v)))
        return np.float32(v.value)

    # -- a6 ---------------------------------------------------------------------------
    def triangulate(self, P1, P2, x1, x2) -> np.ndarray:
        P1 = np.ascontiguousarray(P1, np.float32).reshape(12)
        P2 = np.ascontiguousarray(P2, np.float32).reshape(12)
        x1 = np.ascontiguousarray(x1, np.float32).reshape(-1, 2)
        x2 = np.ascontiguousarray(x2, np.float32).reshape(-1, 2)
        n = len(x1)
        X4 = np.zeros((4, n), np.float32)
        self._check(self._L.vsf_triangulate(self._h, _ptr(P1), _ptr(P2), _ptr(x1), _ptr(x2), n,
                                            _ptr(X4)))
        return X4

    def undistort_points(self, K, dist, xy) -> np.ndarray:
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        d = np.zeros(5, np.float32)
        dd = np.asarray(dist, np.float32).ravel()
        d[:min(5, len(dd))] = dd[:5]
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros_like(xy)
        self._check(self._L.vsf_undistort_points(self._h, _ptr(K), _ptr(d), _ptr(xy), len(xy), _ptr(out)))
        return out

    # -- fused ObserveImage matching path ------------------------------------------------
    def observe_features(self, frame_id: int, kp_left, desc_left, kp_right, desc_right, F,
                         P_left, P_right, ratio: float) -> dict:
        kl = np.ascontiguousarray(kp_left, KEYPOINT_DTYPE)
        kr = np.ascontiguousarray(kp_right, KEYPOINT_DTYPE)
        dl, dr = _u8rows(desc_left, self.desc_bytes), _u8rows(desc_right, self.desc_bytes)
        F = np.ascontiguousarray(F, np.float32).reshape(9)
        P1 = np.ascontiguousarray(P_left, np.float32).reshape(12)
        P2 = np.ascontiguousarray(P_right, np.float32).reshape(12)
        cap = self.max_features          # covers this frame and every resident past frame
        W = self.window
        kept_l, kept_r = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        fids, wc = np.zeros(W, np.uint64), np.zeros(W, np.int32)
        wm = np.zeros((W, cap), DMATCH_DTYPE)
        tm = np.zeros(cap, DMATCH_DTYPE)
        X4 = np.zeros((cap, 4), np.float32)
        o = _ObserveOut(_ptr(kept_l), _ptr(kept_r), 0, 0.0, 0, _ptr(fids), _ptr(wc), _ptr(wm),
                        0, _ptr(tm), _ptr(X4), cap, None)
        self._check(self._L.vsf_observe_features(
            self._h, frame_id, _ptr(kl), _ptr(dl), len(kl), dl.strides[0], _ptr(kr), _ptr(dr),
            len(kr), dr.strides[0], _ptr(F), _ptr(P1), _ptr(P2), float(ratio), C.byref(o)))
        return dict(
            kept_left=kept_l[:o.n_kept].copy(), kept_right=kept_r[:o.n_kept].copy(),
            stereo_threshold_next=np.float32(o.stereo_threshold_next),
            window=[(int(fids[j]), wm[j, :wc[j]].copy()) for j in range(o.n_frames)],
            tri_matches=tm[:o.n_tri].copy(), tri_X4=X4[:o.n_tri].copy())

    # pipelined form: submit whole frames ahead, collect them in order
    def observe_submit(self, frame_id: int, kp_left, desc_left, kp_right, desc_right, F, P_left,
                       P_right, K_left, dist_left, ratio: float):
        """K_left / dist_left may be None (no undistorted pixels are produced then)."""
        kl = np.ascontiguousarray(kp_left, KEYPOINT_DTYPE)
        kr = np.ascontiguousarray(kp_right, KEYPOINT_DTYPE)
        dl, dr = _u8rows(desc_left, self.desc_bytes), _u8rows(desc_right, self.desc_bytes)
        arrs = [np.ascontiguousarray(F, np.float32).reshape(9),
                np.ascontiguousarray(P_left, np.float32).reshape(12),
                np.ascontiguousarray(P_right, np.float32).reshape(12),
                None if K_left is None else np.ascontiguousarray(K_left, np.float32).reshape(9),
                None if dist_left is None else np.ascontiguousarray(dist_left, np.float32).reshape(5)]
        prm = _ObserveParams(*[_ptr(a) for a in arrs], float(ratio))
        self._check(self._L.vsf_observe_submit(
            self._h, frame_id, _ptr(kl), _ptr(dl), len(kl), dl.strides[0], _ptr(kr), _ptr(dr),
            len(kr), dr.strides[0], C.byref(prm)))

    def observe_in_flight(self) -> int:
        return self._L.vsf_observe_in_flight(self._h)

    def observe_collect(self) -> dict:
        cap = self.max_features
        W = self.window
        kept_l, kept_r = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        fids, wc = np.zeros(W, np.uint64), np.zeros(W, np.int32)
        wm = np.zeros((W, cap), DMATCH_DTYPE)
        tm = np.zeros(cap, DMATCH_DTYPE)
        X4 = np.zeros((cap, 4), np.float32)
        xyu = np.full((cap, 2), np.nan, np.float32)
        o = _ObserveOut(_ptr(kept_l), _ptr(kept_r), 0, 0.0, 0, _ptr(fids), _ptr(wc), _ptr(wm),
                        0, _ptr(tm), _ptr(X4), cap, _ptr(xyu))
        fid = C.c_uint64(0)
        self._check(self._L.vsf_observe_collect(self._h, C.byref(fid), C.byref(o)))
        return dict(
            frame_id=int(fid.value),
            kept_left=kept_l[:o.n_kept].copy(), kept_right=kept_r[:o.n_kept].copy(),
            stereo_threshold_next=np.float32(o.stereo_threshold_next),
            window=[(int(fids[j]), wm[j, :wc[j]].copy()) for j in range(o.n_frames)],
            tri_matches=tm[:o.n_tri].copy(), tri_X4=X4[:o.n_tri].copy(),
            xy_undist=xyu[:o.n_kept].copy())

    # -- device-resident entry points -------------------------------------------------------
    def window_match_device(self, d_queries: Sequence[int], nq: Sequence[int], d_train: int,
                            nt: int, ratio: float):
        n = len(d_queries)
        qp = (C.c_void_p * max(n, 1))(*[C.c_void_p(int(p)) for p in d_queries])
        nn = (C.c_int * max(n, 1))(*[int(v) for v in nq])
        self._check(self._L.vsf_window_match_device(self._h, qp, nn, n, C.c_void_p(int(d_train)),
                                                    int(nt), float(ratio)))

    def window_match_block_device(self, d_seq: int, n: int, n_poses: int, first: int, count: int,
                                  ratio: float):
        self._check(self._L.vsf_window_match_block_device(self._h, C.c_void_p(int(d_seq)), n, n_poses,
                                                          first, count, float(ratio)))

    def window_run_sequence(self, h_seq: np.ndarray, first: int, count: int, ratio: float,
                            best_percent: float, sort_mode: int, lag: int, out: np.ndarray,
                            counts: np.ndarray):
        """h_seq: page-locked (n_poses, n, row_bytes) uint8; out: (ring, window, cap) FEATURE_MATCH;
        counts: (ring, window) int32.  -> (h2d_bytes, d2h_bytes) of the run."""
        a, b = C.c_size_t(0), C.c_size_t(0)
        n_poses, n = h_seq.shape[0], h_seq.shape[1]
        self._check(self._L.vsf_window_run_sequence(
            self._h, h_seq.ctypes.data, n, n_poses, first, count, float(ratio), float(best_percent),
            sort_mode, lag, out.ctypes.data, counts.ctypes.data, out.shape[0], out.shape[2],
            C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def fetch_window(self, n_frames: int, with_matches: bool = True):
        counts = np.zeros(max(n_frames, 1), np.int32)
        cap = self.max_features
        out = np.zeros((max(n_frames, 1), cap), DMATCH_DTYPE) if with_matches else None
        self._check(self._L.vsf_fetch_window(self._h, n_frames, _ptr(counts), _ptr(out), cap))
        if not with_matches:
            return counts[:n_frames]
        return [out[j, :counts[j]].copy() for j in range(n_frames)]

    def synth_sequence_device(self, d_out: int, n: int, first_pose: int, n_poses: int,
                              stride: int, seed: int):
        self._check(self._L.vsf_synth_sequence_device(self._h, C.c_void_p(int(d_out)), n,
                                                      first_pose, n_poses, stride, seed))

    def debug_sort_device(self, matches: np.ndarray, best_percent: float, exact: bool) -> np.ndarray:
        """Device sort + cut of a DMATCH list -> (keep, 2) uint64 [initial, current]."""
        m = np.ascontiguousarray(matches, DMATCH_DTYPE)
        out = np.zeros(max(len(m), 1), FEATURE_MATCH_DTYPE)
        n = C.c_int(0)
        self._check(self._L.vsf_debug_sort_device(self._h, _ptr(m), len(m), float(best_percent), int(exact),
                                                  _ptr(out), C.byref(n)))
        fm = out[:n.value]
        return np.stack([fm["feature_idx_initial"], fm["feature_idx_current"]], 1).reshape(-1, 2)

    def probe_pipe(self, kind: int, iters: int = 4096) -> float:
        v = C.c_double(0)
        self._check(self._L.vsf_probe_pipe(self._h, kind, iters, C.byref(v)))
        return v.value
