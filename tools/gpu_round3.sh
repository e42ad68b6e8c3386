#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json;d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], d['roofline_int']['frac'], d['e2e']['value'], d['e2e']['device_sort_variant'], d['cpu_baseline']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
tail -5 gpurun_out/launches.csv
