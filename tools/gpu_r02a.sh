#!/bin/bash
# Round 2, first GPU call: new parity tests, whole GPU suite, smoke, bench (both arms), launch list.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r02a_gpu.txt 2>&1
python -m pytest tests/test_gpu_round2.py -x -q -m gpu > $O/r02a_gpu_tests_new.txt 2>&1; tail -15 $O/r02a_gpu_tests_new.txt
python -m pytest tests -q -m gpu > $O/r02a_gpu_tests.txt 2>&1; tail -8 $O/r02a_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02a_smoke.txt 2>&1; tail -2 $O/r02a_smoke.txt
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02a_bench_reference_arm.json 2> $O/r02a_bench_reference_arm.err; tail -c 300 $O/r02a_bench_reference_arm.json
python bench.py --steps 20 --warmup 5 > $O/r02a_bench_c4.json 2> $O/r02a_bench_c4.err; tail -c 400 $O/r02a_bench_c4.err
python bench.py --steps 20 --warmup 5 --desc-bytes 61 --no-extra > $O/r02a_bench_c4_61.json 2> $O/r02a_bench_c4_61.err; tail -c 400 $O/r02a_bench_c4_61.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r02a_ncu_launches_bench_c4.csv \
    python bench.py --steps 1 --warmup 1 --poses-per-step 64 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_bench.log 2>&1
ls -la $O | tail -20
