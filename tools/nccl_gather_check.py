"""Run under torchrun (one rank per GPU): every rank matches its own pose of the synthetic
sequence against a 3-frame window, vsf_gather_matches all-gathers the device match lists, and
every rank checks every rank's lists against the oracle.  Used by
tests/test_gpu_round2.py::test_gather_matches_two_ranks_over_nccl and tools/gpu_r02_multi.sh."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist

import synth
import vision_slam_frontend_b200 as vsf
from oracle import native
from vision_slam_frontend_b200 import capi, nccl

RATIO = float(np.float32(0.6))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, W, stride, seed = 1500, 3, 150, 31
    comm = nccl.Comm.from_torch_distributed(local)
    with vsf.Context(device=local, max_features=n, window=W) as ctx:
        first = 10 * rank                         # this rank's pose: first + W against first .. first + W - 1
        seq = torch.empty((W + 1, n, 32), dtype=torch.uint8, device="cuda")
        ctx.synth_sequence_device(seq.data_ptr(), n, first, W + 1, stride, seed)
        ctx.window_match_block_device(seq.data_ptr(), n, W + 1, 0, 1, RATIO)
        S = nccl.match_list_stride(ctx)
        d_counts = torch.zeros((world, W), dtype=torch.int32, device="cuda")
        d_lists = torch.zeros((world, W, S, 16), dtype=torch.uint8, device="cuda")
        comm.gather_matches(ctx, W, d_counts.data_ptr(), d_lists.data_ptr())
        ctx.synchronize()
        lists = nccl.gathered_to_lists(d_counts.cpu().numpy(), d_lists.cpu().numpy())
    total = 0
    for r in range(world):
        for j in range(W):
            exp = native.get_matches(synth.synth_pose(n, 10 * r + j, stride, seed),
                                     synth.synth_pose(n, 10 * r + W, stride, seed), RATIO)
            assert len(lists[r][j]) == len(exp) and (lists[r][j] == exp).all(), (rank, r, j)
            total += len(exp)
    assert total > 500
    # and the byte gather the sequence driver uses
    blobs = comm.gather_bytes(bytes([rank]) * (1000 + 77 * rank), root=0)
    if rank == 0:
        assert [len(b) for b in blobs] == [1000 + 77 * r for r in range(world)]
        assert all(b == bytes([r]) * len(b) for r, b in enumerate(blobs))
    comm.close()
    print("gather ok rank %d: %d matches from %d ranks" % (rank, total, world), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
