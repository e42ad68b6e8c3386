#!/bin/bash
mkdir -p gpurun_out
for e in 2 3; do timeout 300 python tools/tc_probe.py case $e 0 > gpurun_out/tc_parity_e$e.log 2>&1; tail -9 gpurun_out/tc_parity_e$e.log | cut -c1-150; done
for f in 0 2 4; do
  N=5000 W=10 POSES=64 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
  N=20000 W=10 POSES=4 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
done 2>&1 | tee gpurun_out/tc_time.log
N=5000 W=10 POSES=64 ENGINE=3 timeout 120 python tools/tc_time.py
N=20000 W=10 POSES=4 ENGINE=3 timeout 120 python tools/tc_time.py
ENGINE=2 N=5000 W=10 LAUNCHES=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_c4.csv python tools/ncu_target.py > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_tc_c4.csv | cut -d, -f5,15- | tail -7
VSF_ENGINE=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
