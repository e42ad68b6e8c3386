#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "device_exact_sort or run_sequence" > $O/r02d_gpu_tests_sort.txt 2>&1; tail -15 $O/r02d_gpu_tests_sort.txt
timeout 900 python -m pytest tests -q -m gpu > $O/r02d_gpu_tests.txt 2>&1; tail -8 $O/r02d_gpu_tests.txt
python bench.py --steps 20 --warmup 5 > $O/r02d_bench_c4.json 2> $O/r02d_bench_c4.err; tail -c 400 $O/r02d_bench_c4.err
VSF_HOST_THREADS=4 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02d_bench_c4_4threads.json 2> $O/r02d_bench_c4_4threads.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 80 --csv --log-file $O/r02d_ncu_launches_probe_c3.csv ./vision_slam_frontend_b200/vsf_latency_probe 0 2000 32 1 > $O/ncu_probe.log 2>&1
ls -la $O | tail -8
