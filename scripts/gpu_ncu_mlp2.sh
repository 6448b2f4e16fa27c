#!/bin/bash
OUT=gpurun_out/${1:-ncu_mlp2}; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp2_ -s 2 -c 1 -o $OUT/prof_mlp2 -f python scripts/mlp2_check.py --profile > $OUT/ncu_mlp2.log 2>&1
tail -2 $OUT/ncu_mlp2.log
