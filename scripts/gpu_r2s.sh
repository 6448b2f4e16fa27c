#!/bin/bash
OUT=gpurun_out/${1:-r2s}; mkdir -p $OUT
for i in 1 2; do
timeout 200 python scripts/mlp2_ab.py | tee -a $OUT/ab.jsonl | cut -c1-400
ALLSET_B200_LIB=$PWD/build/variants/lib_linear.so timeout 200 python scripts/mlp2_ab.py | tee -a $OUT/ab.jsonl | cut -c1-400
done
