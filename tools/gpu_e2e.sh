#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_frontend.py -m gpu -x -q -k "pipelined or window or frontend or observe" 2>&1 | tail -8
timeout 300 python tools/e2e_probe.py 2>&1 | tail -6
for lag in 2 3; do
timeout 600 python bench.py --steps 300 --no-cpu-baseline --e2e-lag $lag > gpurun_out/bench_e2e_lag$lag.json 2> gpurun_out/bench_e2e.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_e2e_lag$lag.json'))
print('lag $lag value', d['value'], 'ms', d['ms_per_step'])
for k, v in d['e2e']['variants'].items(): print('   ', k, v)
PY
done
