#!/bin/bash
mkdir -p gpurun_out
V=vision_slam_frontend_b200/build/variants
for lib in "" $V/b16.so $V/b32.so; do
  echo "=== lib: ${lib:-default(b8)}"
  export VSF_LIB_PATH=$lib
  timeout 300 python tools/tc_probe.py case 2 0 > gpurun_out/tc_parity.log 2>&1; grep -c "'bad_idx_rows': 0, 'bad_dist_rows': 0, 'matches_equal': True" gpurun_out/tc_parity.log; tail -1 gpurun_out/tc_parity.log
  N=5000 W=10 POSES=64 ENGINE=2 timeout 120 python tools/tc_time.py
  N=20000 W=10 POSES=4 ENGINE=2 timeout 120 python tools/tc_time.py
  ENGINE=2 N=5000 W=10 LAUNCHES=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_c4.csv python tools/ncu_target.py > /dev/null 2>&1
  grep -v "^==" gpurun_out/launches_tc_c4.csv | cut -d, -f5,15- | tail -4 | cut -c1-60,120-
done
unset VSF_LIB_PATH
timeout 300 python tools/tc_probe.py case 3 0 > gpurun_out/tc_parity_e3.log 2>&1; grep -c "'bad_idx_rows': 0, 'bad_dist_rows': 0, 'matches_equal': True" gpurun_out/tc_parity_e3.log; tail -1 gpurun_out/tc_parity_e3.log
N=5000 W=10 POSES=64 ENGINE=3 timeout 120 python tools/tc_time.py
N=5000 W=10 POSES=64 ENGINE=2 FLAGS=4 timeout 120 python tools/tc_time.py
N=20000 W=10 POSES=4 ENGINE=2 FLAGS=4 timeout 120 python tools/tc_time.py
VSF_ENGINE=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
