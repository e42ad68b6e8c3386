#!/bin/bash
# Round-2 closing measurement on 8 GPUs of one box: the bench under torchrun, then the reference arm is not needed here.
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02b_bench_c4_8gpu.json 2> $O/r02b_bench_c4_8gpu.err; tail -c 400 $O/r02b_bench_c4_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r02b_bench_c4_4gpu.json 2> $O/r02b_bench_c4_4gpu.err; tail -c 200 $O/r02b_bench_c4_4gpu.err
