#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_gpu.txt
timeout 1200 python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; tail -3 gpurun_out/probe.txt
