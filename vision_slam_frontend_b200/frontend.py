"""ctypes view of libvsf_frontend.so — the C++ slam::Frontend mirror
(csrc/frontend/) — for the tests and for Python callers.  No fallback: raises if
the library is missing or no CUDA device can be opened."""
from __future__ import annotations

import ctypes as C
import os
import struct
from typing import List, Tuple

import numpy as np

from .capi import DMATCH_DTYPE, FEATURE_MATCH_DTYPE, KEYPOINT_DTYPE, load_library

_PKG = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_PKG, "libvsf_frontend.so")
_LIB = None


def load_frontend_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    load_library()                       # libvsf_cuda.so first (rpath $ORIGIN also finds it)
    if not os.path.exists(_PATH):
        raise ImportError(f"{_PATH} not built: run `python -m vision_slam_frontend_b200.build`")
    L = C.CDLL(_PATH)
    vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
    L.vsff_create.argtypes = [i, i, i, i, f, f, vp, vp, vp, vp, vp, i]
    L.vsff_create.restype = vp
    L.vsff_destroy.argtypes = [vp]
    L.vsff_last_error.argtypes = [vp]
    L.vsff_last_error.restype = C.c_char_p
    L.vsff_ok.argtypes = [vp]
    L.vsff_default_config.argtypes = [vp] * 5
    L.vsff_observe_odometry.argtypes = [vp, vp, vp, d]
    L.vsff_observe_features.argtypes = [vp, vp, vp, i, vp, vp, i, i, d]
    L.vsff_get_matches.argtypes = [vp, vp, i, vp, i, i, d, vp, i]
    L.vsff_num_poses.argtypes = [vp]
    L.vsff_stereo_threshold.argtypes = [vp]
    L.vsff_stereo_threshold.restype = f
    L.vsff_num_vision_factors.argtypes = [vp]
    L.vsff_vision_factor.argtypes = [vp, i, vp, vp, vp, i]
    L.vsff_node_features.argtypes = [vp, i, vp, vp, i]
    L.vsff_node_pose.argtypes = [vp, i, vp, vp, vp]
    L.vsff_num_odometry_factors.argtypes = [vp]
    L.vsff_odometry_factor.argtypes = [vp, i, vp, vp, vp, vp]
    L.vsff_serialize_problem.argtypes = [vp, vp, C.c_size_t]
    L.vsff_serialize_problem.restype = C.c_size_t
    L.vsff_submit_features.argtypes = [vp, vp, vp, i, vp, vp, i, i, d]
    L.vsff_collect_features.argtypes = [vp]
    L.vsff_in_flight.argtypes = [vp]
    L.vsff_start_shard.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.vsff_synth_frame.argtypes = [i, i, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vp, C.POINTER(d)]
    L.vsff_run_synthetic_sequence.argtypes = [i, i, i, i, i, i, i, C.c_uint64, i, C.POINTER(vp), C.c_char_p, i]
    L.vsff_run_synthetic_sequence.restype = C.c_size_t
    L.vsff_free.argtypes = [vp]
    L.vsff_synthetic_rig.argtypes = [vp] * 5
    _LIB = L
    return L


def default_config() -> dict:
    L = load_frontend_library()
    P1, P2 = np.zeros(12, np.float32), np.zeros(12, np.float32)
    F, K, D = np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(5, np.float32)
    L.vsff_default_config(P1.ctypes.data, P2.ctypes.data, F.ctypes.data, K.ctypes.data, D.ctypes.data)
    return dict(P_left=P1.reshape(3, 4), P_right=P2.reshape(3, 4), fundamental=F.reshape(3, 3),
                K_left=K.reshape(3, 3), dist_left=D)


def synthetic_rig() -> dict:
    """The rig vsf_sequence_driver / run_synthetic_sequence use (SyntheticRig)."""
    L = load_frontend_library()
    P1, P2 = np.zeros(12, np.float32), np.zeros(12, np.float32)
    F, K, D = np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(5, np.float32)
    L.vsff_synthetic_rig(P1.ctypes.data, P2.ctypes.data, F.ctypes.data, K.ctypes.data, D.ctypes.data)
    return dict(P_left=P1.reshape(3, 4), P_right=P2.reshape(3, 4), fundamental=F.reshape(3, 3),
                K_left=K.reshape(3, 3), dist_left=D)


class Frontend:
    """slam::Frontend (C++) driven from Python."""

    def __init__(self, device=0, max_features=8192, desc_bytes=32, frame_life=10, best_percent=0.3,
                 nn_match_ratio=0.6, P_left=None, P_right=None, fundamental=None, K_left=None,
                 dist_left=None, exact_std_sort=True):
        self._L = load_frontend_library()
        self._keep = [None if a is None else np.ascontiguousarray(a, np.float32).ravel()
                      for a in (P_left, P_right, fundamental, K_left, dist_left)]
        ptrs = [None if a is None else a.ctypes.data for a in self._keep]
        self._h = self._L.vsff_create(device, max_features, desc_bytes, frame_life, best_percent,
                                      nn_match_ratio, *ptrs, int(exact_std_sort))
        self.desc_bytes, self.max_features = desc_bytes, max_features
        if not self._L.vsff_ok(self._h):
            msg = self._L.vsff_last_error(self._h).decode()
            self._L.vsff_destroy(self._h)
            self._h = None
            raise RuntimeError(msg)

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsff_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _desc(self, a) -> np.ndarray:
        """Descriptor matrix as the C++ side reads it: 2-D uint8, exactly desc_bytes wide."""
        a = np.ascontiguousarray(a, np.uint8)
        if a.ndim != 2 or a.shape[1] != self.desc_bytes:
            raise ValueError("descriptors must be (n, %d) uint8, got %r" % (self.desc_bytes, a.shape))
        return a

    def observe_odometry(self, translation, rotation_wxyz, timestamp: float):
        t = np.ascontiguousarray(translation, np.float32)
        q = np.ascontiguousarray(rotation_wxyz, np.float32)
        self._L.vsff_observe_odometry(self._h, t.ctypes.data, q.ctypes.data, float(timestamp))

    def observe_features(self, kp_left, desc_left, kp_right, desc_right, time: float = 0.0) -> bool:
        kl = np.ascontiguousarray(kp_left, KEYPOINT_DTYPE)
        kr = np.ascontiguousarray(kp_right, KEYPOINT_DTYPE)
        dl = self._desc(desc_left)
        dr = self._desc(desc_right)
        if len(dl) != len(kl) or len(dr) != len(kr):
            raise ValueError("keypoint / descriptor row mismatch")
        rc = self._L.vsff_observe_features(self._h, kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data,
                                           dr.ctypes.data, len(kr), self.desc_bytes, float(time))
        if rc < 0:
            raise RuntimeError(self._L.vsff_last_error(self._h).decode())
        return bool(rc)

    # pipelined form (Frontend::SubmitFeatures / CollectFeatures)
    def submit_features(self, kp_left, desc_left, kp_right, desc_right, time: float = 0.0) -> bool:
        kl = np.ascontiguousarray(kp_left, KEYPOINT_DTYPE)
        kr = np.ascontiguousarray(kp_right, KEYPOINT_DTYPE)
        dl, dr = self._desc(desc_left), self._desc(desc_right)
        rc = self._L.vsff_submit_features(self._h, kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data,
                                          dr.ctypes.data, len(kr), self.desc_bytes, float(time))
        if rc < 0:
            raise RuntimeError(self._L.vsff_last_error(self._h).decode())
        return bool(rc)

    def collect_features(self) -> bool:
        rc = self._L.vsff_collect_features(self._h)
        if rc < 0:
            raise RuntimeError(self._L.vsff_last_error(self._h).decode())
        return bool(rc)

    @property
    def in_flight(self) -> int:
        return self._L.vsff_in_flight(self._h)

    def start_shard(self, halo_first: int, first: int):
        if self._L.vsff_start_shard(self._h, halo_first, first) < 0:
            raise RuntimeError(self._L.vsff_last_error(self._h).decode())

    def get_matches(self, Q, T, ratio: float) -> np.ndarray:
        Q, T = self._desc(Q), self._desc(T)
        out = np.zeros(max(len(Q), 1), DMATCH_DTYPE)
        n = self._L.vsff_get_matches(self._h, Q.ctypes.data, len(Q), T.ctypes.data, len(T), self.desc_bytes,
                                     float(ratio), out.ctypes.data, len(out))
        if n < 0:
            raise RuntimeError(self._L.vsff_last_error(self._h).decode())
        return out[:n].copy()

    @property
    def num_poses(self) -> int:
        return self._L.vsff_num_poses(self._h)

    @property
    def stereo_threshold(self) -> np.float32:
        return np.float32(self._L.vsff_stereo_threshold(self._h))

    def vision_factors(self) -> List[Tuple[int, int, np.ndarray]]:
        out = []
        buf = np.zeros(self.max_features, FEATURE_MATCH_DTYPE)
        for i in range(self._L.vsff_num_vision_factors(self._h)):
            a, b = C.c_uint64(0), C.c_uint64(0)
            n = self._L.vsff_vision_factor(self._h, i, C.byref(a), C.byref(b), buf.ctypes.data, len(buf))
            assert n >= 0
            pairs = np.stack([buf["feature_idx_initial"][:n], buf["feature_idx_current"][:n]], 1)
            out.append((int(a.value), int(b.value), pairs.copy()))
        return out

    def node(self, i: int) -> dict:
        px = np.zeros((self.max_features, 2), np.float32)
        p3 = np.zeros((self.max_features, 3), np.float32)
        n = self._L.vsff_node_features(self._h, i, px.ctypes.data, p3.ctypes.data, self.max_features)
        assert n >= 0
        loc, q = np.zeros(3, np.float32), np.zeros(4, np.float32)
        ts = C.c_double(0)
        self._L.vsff_node_pose(self._h, i, loc.ctypes.data, q.ctypes.data, C.byref(ts))
        return dict(pixel=px[:n].copy(), point3d=p3[:n].copy(), loc=loc, quat_xyzw=q, timestamp=ts.value)

    def odometry_factors(self):
        out = []
        for i in range(self._L.vsff_num_odometry_factors(self._h)):
            a, b = C.c_uint64(0), C.c_uint64(0)
            t, q = np.zeros(3, np.float32), np.zeros(4, np.float32)
            self._L.vsff_odometry_factor(self._h, i, C.byref(a), C.byref(b), t.ctypes.data, q.ctypes.data)
            out.append((int(a.value), int(b.value), t, q))
        return out

    def serialize_problem(self) -> bytes:
        n = self._L.vsff_serialize_problem(self._h, None, 0)
        buf = (C.c_uint8 * max(n, 1))()
        self._L.vsff_serialize_problem(self._h, buf, n)
        return bytes(buf[:n])


def synth_frame(features: int, desc_bytes: int, seed: int, pose: int):
    """One frame of the C++ synthetic stereo source (csrc/frontend/synthetic_source.h) and its
    odometry message: (kp_left, desc_left, kp_right, desc_right, translation, quat_wxyz, timestamp)."""
    L = load_frontend_library()
    kl, kr = np.zeros(features, KEYPOINT_DTYPE), np.zeros(features, KEYPOINT_DTYPE)
    dl, dr = np.zeros((features, desc_bytes), np.uint8), np.zeros((features, desc_bytes), np.uint8)
    t, q = np.zeros(3, np.float32), np.zeros(4, np.float32)
    ts = C.c_double(0)
    rc = L.vsff_synth_frame(features, desc_bytes, seed, pose, kl.ctypes.data, dl.ctypes.data, kr.ctypes.data,
                            dr.ctypes.data, t.ctypes.data, q.ctypes.data, C.byref(ts))
    if rc:
        raise RuntimeError("vsff_synth_frame failed")
    return kl, dl, kr, dr, t, q, ts.value


def run_synthetic_sequence(features: int, desc_bytes: int, frame_life: int, n_poses: int, world: int = 1,
                           in_flight: int = 1, seed: int = 1, exact: bool = True, device: int = 0) -> bytes:
    """The SLAMProblem message (ROS1 wire bytes) of poses [0, n_poses) of the synthetic sequence,
    run as `world` shards one after the other in this process and merged - what the ranks of
    vsf_sequence_driver do, minus the NCCL transport of the pieces."""
    L = load_frontend_library()
    out = C.c_void_p()
    err = C.create_string_buffer(512)
    n = L.vsff_run_synthetic_sequence(device, features, desc_bytes, frame_life, n_poses, world, in_flight, seed,
                                      int(exact), C.byref(out), err, 512)
    if n == 0:
        raise RuntimeError(err.value.decode() or "vsff_run_synthetic_sequence failed")
    data = C.string_at(out, n)
    L.vsff_free(out)
    return data


def parse_slam_problem(wire: bytes) -> dict:
    """Decode the ROS1 wire format of vision_slam_frontend/SLAMProblem (msg/*.msg)."""
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from("<" + fmt, wire, off)
        off += struct.calcsize("<" + fmt)
        return v

    nodes = []
    for _ in range(take("I")[0]):
        idx, ts = take("Qd")
        loc = take("3d")
        quat = take("4d")
        feats = []
        for _ in range(take("I")[0]):
            fid, = take("Q")
            pix = take("3d")
            p3 = take("3d")
            feats.append((fid, pix, p3))
        nodes.append(dict(id=idx, timestamp=ts, loc=loc, quat_xyzw=quat, features=feats))
    vfs = []
    for _ in range(take("I")[0]):
        a, b = take("QQ")
        n, = take("I")
        pairs = np.frombuffer(wire, "<u8", 2 * n, off).reshape(n, 2).copy()
        off += 16 * n
        vfs.append((a, b, pairs))
    ofs = []
    for _ in range(take("I")[0]):
        i, j = take("QQ")
        ofs.append((i, j, take("3d"), take("4d")))
    assert off == len(wire), "trailing bytes"
    return dict(nodes=nodes, vision_factors=vfs, odometry_factors=ofs)
