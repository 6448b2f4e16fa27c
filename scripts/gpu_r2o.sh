#!/bin/bash
OUT=gpurun_out/${1:-r2o}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short > $OUT/pytest_linear.txt 2>&1; tail -3 $OUT/pytest_linear.txt
timeout 300 python scripts/linear_bench.py 2>&1 | tee $OUT/linear_bench.jsonl | cut -c1-330
timeout 200 python scripts/mlp2_check.py --time 2>&1 | tee $OUT/mlp2_check.txt | grep -i "timing\|FAIL\|ok" | cut -c1-200 | tail -14
