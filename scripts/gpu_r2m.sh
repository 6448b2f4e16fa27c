#!/bin/bash
OUT=gpurun_out/${1:-r2m}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short > $OUT/pytest_linear.txt 2>&1; tail -5 $OUT/pytest_linear.txt
timeout 300 python scripts/linear_bench.py 2>&1 | tee $OUT/linear_bench.jsonl
