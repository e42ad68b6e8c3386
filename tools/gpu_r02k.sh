#!/bin/bash
set -u
O=gpurun_out
for r in 4 6 8 10 14; do
  VSF_RESERVE_SMS=$r VSF_HOST_THREADS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O/r02k_reserve_$r.json 2> $O/r02k_reserve_$r.err
done
