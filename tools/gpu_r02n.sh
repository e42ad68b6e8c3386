#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:knn2_tc_finish -s 40 -c 2 -o $O/r02b_finish_c4 -f python tools/chain_probe.py 5000 10 32 $O/ncu_probe.json > $O/ncu_full_fin.log 2>&1
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:knn2_tc_finish -s 40 -c 2 -o $O/r02b_finish_c4_61 -f python tools/chain_probe.py 5000 10 61 $O/ncu_probe61.json > $O/ncu_full_fin61.log 2>&1
for f in r02b_finish_c4 r02b_finish_c4_61; do ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null; done
ls -la $O/r02b_finish*
