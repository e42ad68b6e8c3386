"""Multi-GPU plumbing: pose-range sharding and the match-list gather.

Every (past frame, current frame) pair is independent, so a sequence is split
into contiguous pose ranges, one per rank, each rank re-reading a halo of the
`window` frames before its range (plus the 1-float stereo threshold halo, see
vsf_set_stereo_threshold).  There is no collective on the data path; the only
exchange is gathering the variable-length per-rank match lists at the end
(NCCL over NVLink on GPUs, gloo in the CPU tests): one all_gather of the record
counts, one all_gather of the padded record buffers.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

# one gathered record = cv::DMatch + the frame pair it belongs to, as int32[6]:
# pose_initial, pose_current, queryIdx, trainIdx, imgIdx, distance
RECORD_I32 = 6


def pose_range(rank: int, world: int, n_poses: int) -> Tuple[int, int]:
    """Contiguous, balanced [first, last) pose range of `rank`."""
    if not (0 <= rank < world) or n_poses < 0:
        raise ValueError("bad rank/world/n_poses")
    base, rem = divmod(n_poses, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def halo_range(first: int, window: int) -> Tuple[int, int]:
    """Poses a rank must additionally read (not own) to match its first poses."""
    return max(0, first - window), first


def observe_halo_start(first: int, window: int) -> int:
    """First pose a rank must run through the full ObserveImage path (stereo stage included)
    before its own range starts at `first`: the `window` poses whose compacted left frames are
    resident when pose `first` is matched (src/slam_frontend.cc:424-434), plus ONE pose before
    them that only supplies the adaptive stereo threshold - `stereo_ambig_constraint` after a
    frame is the mean residual of that frame's raw matches + 2 (:392-394) and does not depend
    on the threshold the frame itself was filtered with, so one extra pose re-creates the
    1-float halo without any communication."""
    return max(0, first - window - 1)


def pack_records(lists: Sequence[np.ndarray], pose_initial: Sequence[int] = None,
                 pose_current: Sequence[int] = None) -> np.ndarray:
    """DMATCH arrays (one per frame pair) -> (n, 6) int32 records."""
    rows = []
    for j, m in enumerate(lists):
        r = np.zeros((len(m), RECORD_I32), np.int32)
        r[:, 0] = pose_initial[j] if pose_initial is not None else j
        r[:, 1] = pose_current[j] if pose_current is not None else -1
        r[:, 2], r[:, 3], r[:, 4] = m["queryIdx"], m["trainIdx"], m["imgIdx"]
        r[:, 5] = m["distance"].astype(np.int32)
        rows.append(r)
    return np.concatenate(rows, 0) if rows else np.zeros((0, RECORD_I32), np.int32)


def gather_records(records: np.ndarray, device=None) -> List[np.ndarray]:
    """all_gather of variable-length (n, 6) int32 record arrays -> one array per rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(records)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    buf = torch.zeros((cap, RECORD_I32), dtype=torch.int32, device=dev)
    if len(records):
        buf[:len(records)] = torch.from_numpy(np.ascontiguousarray(records)).to(dev)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o[:c].cpu().numpy() for o, c in zip(outs, counts)]


def gather_match_lists(lists: Sequence[np.ndarray], device=None, pose_initial=None,
                       pose_current=None) -> List[List[np.ndarray]]:
    """Gather every rank's per-frame-pair match lists; returns, per rank, the list
    of (n_j, 6) record arrays split back by frame pair (pose_initial)."""
    recs = pack_records(lists, pose_initial, pose_current)
    out = []
    for r in gather_records(recs, device):
        keys = r[:, 0] if len(r) else np.zeros(0, np.int32)
        out.append([r[keys == k] for k in np.unique(keys)] if len(r) else [])
    return out
