#!/bin/bash
# round 2, call L (1 GPU): ncu --set full of the tcgen05 Linear kernels; the pages are exported on the box (the report is > 64 MiB)
OUT=gpurun_out/${1:-r2l}; mkdir -p $OUT
TMP=/tmp/prof_linear
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp2_ws_kernel|wgrad_kernel' -c 4 -o $TMP python scripts/prof_linear.py > $OUT/ncu_linear.log 2>&1; tail -3 $OUT/ncu_linear.log
ncu -i $TMP.ncu-rep --page raw --csv > $OUT/linear_raw.csv 2>/dev/null
ncu -i $TMP.ncu-rep --page source --csv --kernel-name regex:mlp2_ws_kernel > $OUT/linear_fwd_source.csv 2>/dev/null
ncu -i $TMP.ncu-rep --page source --csv --kernel-name regex:wgrad_kernel > $OUT/linear_wgrad_source.csv 2>/dev/null
ls -la $OUT
