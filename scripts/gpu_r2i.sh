#!/bin/bash
# round 2, call I (1 GPU): the tcgen05 Linear kernels (split-precision forward / dgrad, MN-major wgrad), then the training path
OUT=gpurun_out/${1:-r2i}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short > $OUT/pytest_linear.txt 2>&1; tail -40 $OUT/pytest_linear.txt
echo "== wgrad with LBO/SBO exchanged (debug)"
ALLSET_WGRAD_SWAP=1 timeout 300 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=line -k "wgrad and not 300001" > $OUT/pytest_wgrad_swap.txt 2>&1; tail -5 $OUT/pytest_wgrad_swap.txt
echo "== training path + model tests"
timeout 900 python -m pytest tests/test_train_path.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short \
  -k "train or setgnn or fused_dense or mlp or layer" > $OUT/pytest_train.txt 2>&1; tail -15 $OUT/pytest_train.txt
timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-250
