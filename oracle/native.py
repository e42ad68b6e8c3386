"""ctypes access to the plain-C kNN oracle (oracle/oracle_knn.c).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import build as _build
from .restate import DMATCH_DTYPE

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(_build.ensure_built()["knn"])
        L.oracle_knn2_hamming.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
        L.oracle_knn2_hamming.restype = None
        L.oracle_get_matches.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p]
        L.oracle_get_matches.restype = ctypes.c_int
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def num_threads() -> int:
    return _lib().oracle_num_threads()


def knn2_hamming(Q: np.ndarray, T: np.ndarray):
    Q = np.ascontiguousarray(Q, dtype=np.uint8)
    T = np.ascontiguousarray(T, dtype=np.uint8)
    nq, nt = len(Q), len(T)
    bytes_ = Q.shape[1] if nq else (T.shape[1] if nt else 32)
    idx = np.full((nq, 2), -1, dtype=np.int32)
    dist = np.full((nq, 2), -1, dtype=np.int32)
    if nq:
        _lib().oracle_knn2_hamming(Q.ctypes.data, nq, T.ctypes.data, nt, bytes_,
                                   idx.ctypes.data, dist.ctypes.data)
    return idx, dist


def get_matches(Q: np.ndarray, T: np.ndarray, ratio: float) -> np.ndarray:
    Q = np.ascontiguousarray(Q, dtype=np.uint8)
    T = np.ascontiguousarray(T, dtype=np.uint8)
    nq, nt = len(Q), len(T)
    if nq == 0 or nt < 2:
        return np.zeros(0, dtype=DMATCH_DTYPE)
    raw = np.zeros((nq, 4), dtype=np.int32)
    si = np.empty((nq, 2), dtype=np.int32)
    sd = np.empty((nq, 2), dtype=np.int32)
    n = _lib().oracle_get_matches(Q.ctypes.data, nq, T.ctypes.data, nt,
                                  Q.shape[1], float(ratio), raw.ctypes.data,
                                  si.ctypes.data, sd.ctypes.data)
    out = np.zeros(n, dtype=DMATCH_DTYPE)
    out["queryIdx"] = raw[:n, 0]
    out["trainIdx"] = raw[:n, 1]
    out["distance"] = raw[:n, 3].astype(np.float32)
    return out
