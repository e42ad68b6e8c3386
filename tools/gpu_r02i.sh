#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/r02i_gpu_tests.txt 2>&1; echo "tests rc=$?"; tail -5 $O/r02i_gpu_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --desc-bytes 61 --no-extra > $O/r02i_bench_c4_61.json 2> $O/r02i_bench_c4_61.err; echo "bench61 rc=$?"; tail -c 300 $O/r02i_bench_c4_61.err
