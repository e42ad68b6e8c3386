#!/bin/bash
# full round check: GPU tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'engine', d.get('engine'))
print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','kernel_ms')})
print('kernel_ms', d['kernel_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['device_sort_variant'])
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('clocks', d['clocks'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
PY
