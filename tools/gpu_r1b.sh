#!/bin/bash
# round-1 re-entry check: pipelined submit/collect parity, full GPU suite, smoke, both bench arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined" 2>&1 | tail -15 | tee gpurun_out/pytest_pipelined.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
for lag in 1 3; do
  timeout 600 python bench.py --steps 300 --no-cpu-baseline --e2e-lag $lag > gpurun_out/bench_lag$lag.json 2>> gpurun_out/bench.err
done
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'engine', d.get('engine'))
print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','kernel_ms')})
print('kernel_ms', d['kernel_ms'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
for k, v in d['e2e']['variants'].items(): print('   ', k, v)
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print('clocks', d['clocks'])
r=json.load(open('gpurun_out/bench_ref.json')); print('ref', r['value'], r['ms_per_step'])
for lag in (1, 3):
    try:
        e=json.load(open('gpurun_out/bench_lag%d.json' % lag))['e2e']
        print('lag', lag, e['value'], e['ms_per_step'], e['variants']['pipelined_device_sort'])
    except Exception as ex: print('lag', lag, 'failed', ex)
PY
