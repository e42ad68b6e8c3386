#!/bin/bash
# Round-1 closing measurements on one B200 (run under gpurun from the repo root).
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/r01c_gpu_tests.txt 2>&1; tail -2 $O/r01c_gpu_tests.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/r01c_bench_reference_arm.json 2> $O/r01c_bench_reference_arm.err; tail -c 300 $O/r01c_bench_reference_arm.json
python bench.py --steps 1000 --warmup 20 > $O/r01c_bench_c4.json 2> $O/r01c_bench_c4.err; tail -c 200 $O/r01c_bench_c4.err
python bench.py --steps 1000 --warmup 20 --engine 1 --no-cpu-baseline > $O/r01c_bench_c4_popc.json 2> $O/r01c_bench_c4_popc.err
# launch list of the same bench command (short), per-launch durations
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file $O/r01c_ncu_launches_bench_c4.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_bench.log 2>&1
# full capture of the four kernels of one launch sequence
ENGINE=2 LAUNCHES=6 ncu --set full --clock-control none --import-source on -k regex:'knn2_tc|expand_train|knn2_compact' -s 12 -c 4 \
    -o $O/r01c_tensor_engine_c4 -f python tools/ncu_target.py > $O/ncu_full.log 2>&1
ncu -i $O/r01c_tensor_engine_c4.ncu-rep --page raw --csv > $O/r01c_ncu_full_raw.csv 2>/dev/null
python tools/pose_timeline.py 5000 10 $O/r01c_pose_timeline_c4.json > /dev/null 2>&1
if [ -d _ab/trace ]; then (cd _ab/trace && python tools/tc_timeline.py 5000 10 ../../$O/r01c_tc_timeline_c4.json > /dev/null 2>&1); fi
python tools/tc_scaling.py $O/r01c_tc_scaling.json > /dev/null 2>&1
python tools/bench_configs.py > $O/r01c_configs.log 2>&1; cp $O/configs.json $O/r01c_configs_c1_c2_c3_c5.json 2>/dev/null
ls -la $O | tail -20
python tools/engine_crossover.py $O/r01c_engine_crossover.json > /dev/null 2>&1
