#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_probe.py case 2 0 > gpurun_out/tc_parity.log 2>&1; tail -9 gpurun_out/tc_parity.log | cut -c1-160
for f in 0 2 4; do
  N=5000 W=10 POSES=64 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
  N=20000 W=10 POSES=4 ENGINE=2 FLAGS=$f timeout 120 python tools/tc_time.py
done 2>&1 | tee gpurun_out/tc_time.log
ENGINE=2 N=5000 W=10 LAUNCHES=4 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_c4.csv python tools/ncu_target.py > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_tc_c4.csv | cut -d, -f5,12- | tail -14
ENGINE=2 N=20000 W=10 LAUNCHES=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_tc_n20000 python tools/ncu_target.py > gpurun_out/ncu_n20000.log 2>&1; tail -2 gpurun_out/ncu_n20000.log
ENGINE=2 N=5000 W=10 LAUNCHES=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2_tc -s 2 -c 2 -f -o gpurun_out/prof_tc_c4 python tools/ncu_target.py > gpurun_out/ncu_c4.log 2>&1; tail -2 gpurun_out/ncu_c4.log
