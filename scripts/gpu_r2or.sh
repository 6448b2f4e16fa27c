#!/bin/bash
OUT=gpurun_out/${1:-r2or}; mkdir -p $OUT
timeout 200 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 100 --tb=short -k "oracle_restatement" 2>&1 | tail -6 | tee $OUT/pytest.txt
