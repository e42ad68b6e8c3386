#!/bin/bash
# Round-1 state check: GPU tests, smoke, both bench arms, ncu launch list of the bench command,
# one ncu --set full capture of the tensor-core kernel + refine kernel (C4 shape).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tee gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --engine 1 --no-cpu-baseline > gpurun_out/bench_popc.json 2> gpurun_out/bench_popc.err
python - <<'PY'
import json
for f in ('gpurun_out/bench.json','gpurun_out/bench_popc.json'):
    d=json.load(open(f))
    print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'engine', d.get('engine'))
    print(' roofline', {k: d['roofline'].get(k) for k in ('bound','achieved','peak','frac','kernel_ms')})
    print(' kernel_ms', d['kernel_ms'])
    print(' e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['device_sort_variant'])
    print(' clocks', d['clocks'])
    if 'cpu_baseline' in d: print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_times.py gpurun_out/launches_bench.csv 3
ENGINE=2 N=5000 W=10 LAUNCHES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn2_tc_kernel|knn2_tc_refine|knn2_compact|expand_train' -s 4 -c 4 -f -o gpurun_out/r01_tc_full python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
