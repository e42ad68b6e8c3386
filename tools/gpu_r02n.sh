#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu  > $O/r02n_tests.txt 2>&1; tail -8 $O/r02n_tests.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02n_bench_c4.json 2> $O/r02n_bench_c4.err; tail -c 400 $O/r02n_bench_c4.err
