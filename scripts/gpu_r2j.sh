#!/bin/bash
OUT=gpurun_out/${1:-r2j}; mkdir -p $OUT
timeout 300 python scripts/diag_split.py citeseer_allsettransformer.pt 2>&1 | tee $OUT/diag_citeseer_f32.txt | tail -8
timeout 300 python scripts/diag_split.py cora_alldeepsets.pt 2>&1 | tee $OUT/diag_cora_f32.txt | tail -8
timeout 300 python scripts/diag_split.py cora_alldeepsets.pt bf16 2>&1 | tee $OUT/diag_cora_bf16.txt | tail -8
timeout 300 python scripts/diag_split.py citeseer_allsettransformer.pt bf16 2>&1 | tee $OUT/diag_citeseer_bf16.txt | tail -8
