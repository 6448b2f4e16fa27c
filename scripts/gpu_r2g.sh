#!/bin/bash
# round 2, call G (1 GPU): rowop kernels after the hash / transposed-reduction change, training-step time
OUT=gpurun_out/${1:-r2g}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_train_path.py tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short \
  -k "rowop or train or world1 or fused_dense or bias_act or setgnn_real" > $OUT/pytest.txt 2>&1; tail -15 $OUT/pytest.txt
timeout 600 python scripts/prof_train.py 12 > $OUT/prof_train.txt 2>&1; grep "====" $OUT/prof_train.txt
timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-220
