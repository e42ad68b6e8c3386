"""Host-side multi-rank logic on CPU: pose-range partition and the match-list
gather, with torch.distributed's gloo backend at world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vision_slam_frontend_b200 import sharding
from vision_slam_frontend_b200.capi import DMATCH_DTYPE


def test_pose_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 7, 8, 100, 10000, 100001):
            spans = [sharding.pose_range(r, world, n) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sharding.halo_range(0, 10) == (0, 0)
    assert sharding.halo_range(4, 10) == (0, 4)
    assert sharding.halo_range(1250, 32) == (1218, 1250)
    with pytest.raises(ValueError):
        sharding.pose_range(2, 2, 10)


def _fake_lists(rank):
    rng = np.random.default_rng(100 + rank)
    lists = []
    for j in range(3):
        m = np.zeros(int(rng.integers(0, 50)) + rank * 7, DMATCH_DTYPE)
        m["queryIdx"] = np.arange(len(m))
        m["trainIdx"] = rng.integers(0, 5000, len(m))
        m["distance"] = rng.integers(0, 60, len(m)).astype(np.float32)
        lists.append(m)
    return lists


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lists = _fake_lists(rank)
        first, last = sharding.pose_range(rank, world, 11)
        got = sharding.gather_match_lists(lists, pose_initial=[first + j for j in range(3)],
                                          pose_current=[first + 3] * 3)
        ok = len(got) == world
        for r in range(world):
            exp = _fake_lists(r)
            f, _ = sharding.pose_range(r, world, 11)
            non_empty = [(j, m) for j, m in enumerate(exp) if len(m)]
            ok &= len(got[r]) == len(non_empty)
            for rec, (j, m) in zip(got[r], non_empty):
                ok &= bool((rec[:, 0] == f + j).all() and (rec[:, 1] == f + 3).all())
                ok &= bool((rec[:, 2] == m["queryIdx"]).all() and (rec[:, 3] == m["trainIdx"]).all())
                ok &= bool((rec[:, 5] == m["distance"].astype(np.int32)).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_match_lists_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
