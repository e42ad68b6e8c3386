#!/bin/bash
# One GPU-box visit: parity tests, smoke, probe/sweep, bench (both arms).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt | tail -8
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
timeout 900 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; tail -5 gpurun_out/probe.txt
timeout 600 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
cut -c1-1500 gpurun_out/bench.json
