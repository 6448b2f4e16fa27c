#!/bin/bash
OUT=gpurun_out/${1:-misc}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl
