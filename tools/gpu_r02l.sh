#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/r02l_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -4 $O/r02l_gpu_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02l_bench_c4.json 2> $O/r02l_bench_c4.err; tail -c 300 $O/r02l_bench_c4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --desc-bytes 61 > $O/r02l_bench_c4_61.json 2> $O/r02l_bench_c4_61.err
