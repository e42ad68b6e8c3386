#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02n_bench_c4.json 2> $O/r02n_bench_c4.err; tail -c 300 $O/r02n_bench_c4.err
