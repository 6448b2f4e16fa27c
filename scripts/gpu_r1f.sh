#!/bin/bash
OUT=gpurun_out/${1:-r1f}; mkdir -p $OUT
echo "== mlp2_check"; timeout 200 python scripts/mlp2_check.py --time 2>&1 | tail -6 | cut -c1-330 | tee $OUT/mlp2_check.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== model bench"; timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl
