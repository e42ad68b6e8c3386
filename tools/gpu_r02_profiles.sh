#!/bin/bash
# Round-2 ncu evidence: full captures of the hot kernels (one GPU, short targets), raw CSV exports.
set -u
O=gpurun_out
mkdir -p $O
ENGINE=2 LAUNCHES=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'knn2_tc|expand_train|knn2_compact' -s 12 -c 4 \
    -o $O/r02_tensor_engine_c4 -f python tools/ncu_target.py > $O/ncu_full_tc.log 2>&1
WIDTH=61 ENGINE=2 LAUNCHES=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc64' -s 6 -c 4 \
    -o $O/r02_tc64_pair_c4 -f python tools/ncu_target.py > $O/ncu_full_tc64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sort_cut' -s 40 -c 2 \
    -o $O/r02_sort_exact_c4 -f python tools/sort_probe.py > $O/ncu_full_sort.log 2>&1
for f in r02_tensor_engine_c4 r02_tc64_pair_c4 r02_sort_exact_c4; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
ls -la $O/*.ncu-rep
