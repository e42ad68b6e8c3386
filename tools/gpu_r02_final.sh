#!/bin/bash
# Round-2 closing measurements on one B200 (run under gpurun from the repo root).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > $O/r02_gpu_tests.txt 2>&1; tail -3 $O/r02_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.txt 2>&1; tail -1 $O/r02_smoke.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err; tail -c 300 $O/r02_bench_c4.err
timeout 600 python bench.py --steps 20 --warmup 5 --desc-bytes 61 > $O/r02_bench_c4_61.json 2> $O/r02_bench_c4_61.err; tail -c 300 $O/r02_bench_c4_61.err
timeout 600 python bench.py --steps 20 --warmup 5 --engine 1 --no-extra --no-cpu-baseline > $O/r02_bench_c4_popc.json 2> $O/r02_bench_c4_popc.err
VSF_HOST_THREADS=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02_bench_c4_4hostthreads.json 2> $O/r02_bench_c4_4hostthreads.err
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 32 1 > $O/r02_latency_probe_c2_c3.json 2> /dev/null
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 61 10 > $O/r02_latency_probe_61_w10.json 2> /dev/null
python tools/sort_probe.py > $O/r02_sort_probe.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02_ncu_launches_bench_c4.csv \
    python bench.py --steps 1 --warmup 1 --poses-per-step 64 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_target.py > $O/r02_sanitizer_memcheck.txt 2>&1; tail -3 $O/r02_sanitizer_memcheck.txt
ls -la $O | tail -5
