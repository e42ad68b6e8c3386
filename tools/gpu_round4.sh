#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1; tail -6 gpurun_out/configs.log | cut -c1-700
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['roofline_int']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['device_sort_variant'], d['cpu_baseline']['value'])"
