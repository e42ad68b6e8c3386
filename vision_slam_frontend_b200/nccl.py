"""ctypes view of libvsf_nccl.so (include/vsf_nccl.h): the NCCL result gathers of the sharded
path.  The communicator is the library's own (ncclCommInitRank from a unique id); under
torchrun the id travels from rank 0 to the others through torch.distributed."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .capi import DMATCH_DTYPE, load_library

_PKG = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_PKG, "libvsf_nccl.so")
_LIB = None
UNIQUE_ID_BYTES = 128


def load_nccl_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    load_library()
    if not os.path.exists(_PATH):
        raise ImportError(f"{_PATH} not built: run `python -m vision_slam_frontend_b200.build`")
    L = C.CDLL(_PATH)
    vp, i = C.c_void_p, C.c_int
    L.vsf_nccl_unique_id.argtypes = [C.c_char_p]
    L.vsf_nccl_comm_create.argtypes = [C.c_char_p, i, i, i, C.POINTER(vp)]
    L.vsf_nccl_comm_adopt.argtypes = [vp, i, i, i, C.POINTER(vp)]
    L.vsf_nccl_comm_destroy.argtypes = [vp]
    L.vsf_nccl_comm_destroy.restype = None
    L.vsf_nccl_last_error.argtypes = [vp]
    L.vsf_nccl_last_error.restype = C.c_char_p
    L.vsf_gather_matches.argtypes = [vp, vp, i, vp, vp]
    L.vsf_nccl_gather_bytes.argtypes = [vp, vp, C.c_size_t, i, C.POINTER(vp), C.POINTER(C.c_size_t)]
    _LIB = L
    return L


EXPORTED_SYMBOLS = ["vsf_nccl_unique_id", "vsf_nccl_comm_create", "vsf_nccl_comm_adopt", "vsf_nccl_comm_destroy",
                    "vsf_nccl_last_error", "vsf_gather_matches", "vsf_nccl_gather_bytes"]


def unique_id() -> bytes:
    buf = C.create_string_buffer(UNIQUE_ID_BYTES)
    if load_nccl_library().vsf_nccl_unique_id(buf):
        raise RuntimeError("ncclGetUniqueId failed")
    return buf.raw


class Comm:
    def __init__(self, uid: bytes, world: int, rank: int, device: int):
        self._L = load_nccl_library()
        h = C.c_void_p()
        if self._L.vsf_nccl_comm_create(uid, world, rank, device, C.byref(h)):
            raise RuntimeError("ncclCommInitRank failed")
        self._h, self.world, self.rank, self.device = h, world, rank, device

    @classmethod
    def from_torch_distributed(cls, device: int):
        """Rank 0 makes the id, torch.distributed carries it to the other ranks."""
        import torch.distributed as dist
        box = [unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], dist.get_world_size(), dist.get_rank(), device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsf_nccl_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def gather_matches(self, ctx, n_frames: int, d_counts: int, d_lists: int):
        """d_counts: device int32 [world][n_frames]; d_lists: device DMATCH [world][n_frames][stride]
        (raw device pointers).  Asynchronous on the ctx stream."""
        rc = self._L.vsf_gather_matches(ctx._h, self._h, n_frames, C.c_void_p(d_counts), C.c_void_p(d_lists))
        if rc:
            raise RuntimeError(self._L.vsf_nccl_last_error(self._h).decode() or f"vsf_gather_matches: {rc}")

    def gather_bytes(self, data: bytes, root: int = 0):
        """-> list of per-rank byte strings on root, None elsewhere."""
        out = C.c_void_p()
        sizes = (C.c_size_t * self.world)()
        rc = self._L.vsf_nccl_gather_bytes(self._h, data, len(data), root, C.byref(out), sizes)
        if rc:
            raise RuntimeError(self._L.vsf_nccl_last_error(self._h).decode() or f"vsf_nccl_gather_bytes: {rc}")
        if self.rank != root:
            return None
        blob = C.string_at(out, sum(sizes))
        C.CDLL(None).free(out)
        res, off = [], 0
        for s in sizes:
            res.append(blob[off:off + s])
            off += s
        return res


def match_list_stride(ctx) -> int:
    """Records per region of the ctx's device match lists (vsf_device_match_lists)."""
    a, b = C.c_void_p(), C.c_void_p()
    stride, regions = C.c_int(0), C.c_int(0)
    rc = ctx._L.vsf_device_match_lists(ctx._h, C.byref(a), C.byref(b), C.byref(stride), C.byref(regions))
    if rc:
        raise RuntimeError("vsf_device_match_lists failed")
    return stride.value


def gathered_to_lists(counts: np.ndarray, lists: np.ndarray):
    """Host copies of the gathered buffers -> per rank, per frame DMATCH arrays."""
    world, nf = counts.shape
    lists = lists.view(DMATCH_DTYPE).reshape(world, nf, -1)
    return [[lists[r, j, :counts[r, j]].copy() for j in range(nf)] for r in range(world)]
