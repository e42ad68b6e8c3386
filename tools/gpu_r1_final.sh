#!/bin/bash
# round check of the committed state: GPU tests, smoke, both bench arms, ncu launch list of the bench command
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_times.py gpurun_out/launches_bench.csv | tail -12
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log | cut -c1-1500
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'engine', d.get('engine'))
print('roofline', {k: d['roofline'].get(k) for k in ('bound','achieved','peak','frac','kernel_ms','traffic')})
print('kernel_ms', d['kernel_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], json.dumps(d['e2e']['variants']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('clocks', d['clocks'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
