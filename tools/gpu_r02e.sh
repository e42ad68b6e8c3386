#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -x -q -m gpu -k "device_exact_sort or run_sequence or pipelined or feature_matches" > $O/r02e_gpu_tests_sort.txt 2>&1; tail -15 $O/r02e_gpu_tests_sort.txt
python tools/sort_probe.py > $O/r02e_sort_probe.txt 2>&1; cat $O/r02e_sort_probe.txt
ncu -k regex:sort_cut --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02e_sort_probe_ncu.csv python tools/sort_probe.py > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02e_bench_c4.json 2> $O/r02e_bench_c4.err; tail -c 400 $O/r02e_bench_c4.err
VSF_HOST_THREADS=4 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02e_bench_c4_4threads.json 2> $O/r02e_bench_c4_4threads.err
timeout 900 python -m pytest tests -q -m gpu > $O/r02e_gpu_tests.txt 2>&1; tail -8 $O/r02e_gpu_tests.txt
