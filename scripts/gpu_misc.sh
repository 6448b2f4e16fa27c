#!/bin/bash
OUT=gpurun_out/${1:-misc}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
timeout 600 python scripts/model_bench.py --cpu 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl
python scripts/kbench.py 1000000 200000 20 128 2>&1 | tail -1 | tee $OUT/kb_cfg3.json
KB_DTYPE=f32 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kb_cfg4_f32.json
KB_GRAPH=powerlaw python scripts/kbench.py 20000000 3200000 12 256 2>&1 | tail -1 | tee $OUT/kb_cfg5_scaled.json
python scripts/kbench.py 10000000 2000000 30 64 2>&1 | tail -1 | tee $OUT/kb_d64.json
