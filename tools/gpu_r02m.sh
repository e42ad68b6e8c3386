#!/bin/bash
set -u
O=gpurun_out
for r in 10 12; do VSF_SORT_STREAMS=1 VSF_RESERVE_SMS=$r VSF_HOST_THREADS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O/r02m_one_$r.json 2>> $O/r02m.err; done
for r in 16 20; do VSF_RESERVE_SMS=$r VSF_HOST_THREADS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline > $O/r02m_two_$r.json 2>> $O/r02m.err; done
