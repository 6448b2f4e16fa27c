#!/bin/bash
OUT=gpurun_out/${1:-r2r}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short -k "linear or mlp2 or pma_tail or score or tcgen05" > $OUT/pytest_tc.txt 2>&1; tail -3 $OUT/pytest_tc.txt
for i in 1 2; do
timeout 200 python scripts/mlp2_ab.py | tee -a $OUT/ab.jsonl
ALLSET_MLP2_L2_PREFETCH=0 timeout 200 python scripts/mlp2_ab.py | tee -a $OUT/ab.jsonl
done
