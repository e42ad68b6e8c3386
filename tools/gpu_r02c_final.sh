#!/bin/bash
# Round-2 closing measurements (one batch per pose group, re-structured finish kernel) on one B200 (run under gpurun from the repo root).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/r02c_gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu > $O/r02c_gpu_tests.txt 2>&1; tail -3 $O/r02c_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02c_smoke.txt 2>&1; tail -1 $O/r02c_smoke.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02c_bench_reference_arm.json 2> $O/r02c_bench_reference_arm.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02c_bench_c4.json 2> $O/r02c_bench_c4.err; tail -c 300 $O/r02c_bench_c4.err
timeout 600 python bench.py --steps 20 --warmup 5 --desc-bytes 61 > $O/r02c_bench_c4_61.json 2> $O/r02c_bench_c4_61.err; tail -c 300 $O/r02c_bench_c4_61.err
timeout 600 python bench.py --steps 20 --warmup 5 --engine 1 --no-extra --no-cpu-baseline > $O/r02c_bench_c4_popc.json 2> $O/r02c_bench_c4_popc.err
VSF_HOST_THREADS=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline > $O/r02c_bench_c4_4hostthreads.json 2> $O/r02c_bench_c4_4hostthreads.err
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 32 1 > $O/r02c_latency_probe_c2_c3.json 2> /dev/null
./vision_slam_frontend_b200/vsf_latency_probe 0 2000 61 10 > $O/r02c_latency_probe_61_w10.json 2> /dev/null
python tools/sort_probe.py > $O/r02c_sort_probe.txt 2>&1
timeout 300 python tools/chain_probe.py 5000 10 32 $O/r02c_pose_groups_probe_c4.json > /dev/null 2>&1
timeout 300 python tools/chain_probe.py 5000 10 61 $O/r02c_pose_groups_probe_c4_61.json > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02c_ncu_launches_bench_c4.csv \
    python bench.py --steps 1 --warmup 1 --poses-per-step 64 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_bench.log 2>&1
# ncu evidence: full captures of the distance and finish kernels inside a grouped block (serialised by ncu)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn2_tc_kernel|knn2_tc_finish' -s 40 -c 4 \
    -o $O/r02c_tensor_engine_c4 -f python tools/chain_probe.py 5000 10 32 $O/ncu_probe.json > $O/ncu_full_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc64_pair|knn2_tc_finish' -s 40 -c 4 \
    -o $O/r02c_tc64_pair_c4 -f python tools/chain_probe.py 5000 10 61 $O/ncu_probe61.json > $O/ncu_full_tc64.log 2>&1
for f in r02c_tensor_engine_c4 r02c_tc64_pair_c4; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
ls -la $O | tail -5
