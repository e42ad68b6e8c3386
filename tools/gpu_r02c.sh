#!/bin/bash
# Round 2, third GPU call: device replay of std::sort (sort_mode 2), latency probe, bench.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "device_exact_sort or run_sequence" > $O/r02c_gpu_tests_sort.txt 2>&1; tail -15 $O/r02c_gpu_tests_sort.txt
timeout 900 python -m pytest tests -q -m gpu > $O/r02c_gpu_tests.txt 2>&1; tail -8 $O/r02c_gpu_tests.txt
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 32 1 > $O/r02c_latency_probe.json 2> $O/r02c_latency_probe.err; cat $O/r02c_latency_probe.json
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 61 10 > $O/r02c_latency_probe_61_w10.json 2>> $O/r02c_latency_probe.err; cat $O/r02c_latency_probe_61_w10.json
python bench.py --steps 20 --warmup 5 > $O/r02c_bench_c4.json 2> $O/r02c_bench_c4.err; tail -c 400 $O/r02c_bench_c4.err
VSF_HOST_THREADS=4 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02c_bench_c4_4threads.json 2> $O/r02c_bench_c4_4threads.err
ls -la $O | tail -8
