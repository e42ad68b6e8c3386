#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_tensor_engine.py -x -q -m gpu  > $O/r02n_tests.txt 2>&1; tail -3 $O/r02n_tests.txt
timeout 200 python tools/chain_probe.py 5000 10 32 $O/r02n_finish_probe.json > $O/r02n_chain_probe.txt 2>&1; tail -c 4000 $O/r02n_chain_probe.txt | cut -c1-600
