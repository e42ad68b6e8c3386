#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_probe.py case 2 0 > gpurun_out/tc_parity.log 2>&1; grep -c "'bad_idx_rows': 0, 'bad_dist_rows': 0, 'matches_equal': True" gpurun_out/tc_parity.log; tail -1 gpurun_out/tc_parity.log
for f in 0 8; do
  N=5000 W=10 POSES=64 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
  N=20000 W=10 POSES=4 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
done
N=2000 W=1 POSES=100 ENGINE=2 timeout 120 python tools/tc_time.py
N=2000 W=1 POSES=100 ENGINE=1 timeout 120 python tools/tc_time.py
ENGINE=2 FLAGS=8 N=5000 W=10 LAUNCHES=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_c4.csv python tools/ncu_target.py > /dev/null 2>&1
python tools/launch_times.py gpurun_out/launches_tc_c4.csv
ENGINE=2 FLAGS=8 N=20000 W=10 LAUNCHES=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_n20000.csv python tools/ncu_target.py > /dev/null 2>&1
python tools/launch_times.py gpurun_out/launches_tc_n20000.csv
VSF_ENGINE=2 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
