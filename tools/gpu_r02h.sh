#!/bin/bash
# CTA-pair kernel bring-up: every command under its own timeout (a hang must not take the box down)
set -u
O=gpurun_out
mkdir -p $O
timeout 180 python -m pytest tests/test_gpu_tensor_engine.py -x -q -m gpu -k "wide" > $O/r02h_wide_tests.txt 2>&1; echo "wide tests rc=$?"; tail -12 $O/r02h_wide_tests.txt
timeout 120 python tools/wide_descriptors.py $O/r02h_wide_descriptors.json > $O/r02h_wide.log 2>&1; echo "wide_descriptors rc=$?"; tail -4 $O/r02h_wide.log
