#!/bin/bash
# round 2, call K (1 GPU): three-term split Linear, diag, training path, fp32 / bf16 model bench, kernel timings
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_linear_tc.py -m gpu -q -p no:cacheprovider --timeout 120 --tb=short > $OUT/pytest_linear.txt 2>&1; tail -25 $OUT/pytest_linear.txt
timeout 300 python scripts/diag_split.py citeseer_allsettransformer.pt 2>&1 | tee $OUT/diag_citeseer_f32.txt | tail -6
timeout 300 python scripts/diag_split.py cora_alldeepsets.pt 2>&1 | tee $OUT/diag_cora_f32.txt | tail -6
echo "== training path + model tests"
timeout 900 python -m pytest tests/test_train_path.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short \
  -k "train or setgnn or fused_dense or mlp or layer" > $OUT/pytest_train.txt 2>&1; tail -15 $OUT/pytest_train.txt
timeout 300 python scripts/linear_bench.py 2>&1 | tee $OUT/linear_bench.jsonl
timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-250
