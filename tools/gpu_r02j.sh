#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -m gpu -k "block_device or c4_full or device_entry" > $O/r02j_tests.txt 2>&1; tail -3 $O/r02j_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e > $O/r02j_bench_c4.json 2> $O/r02j_bench_c4.err; tail -c 300 $O/r02j_bench_c4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline --no-e2e --desc-bytes 61 > $O/r02j_bench_c4_61.json 2> $O/r02j_bench_c4_61.err
