#!/bin/bash
# round 2, last call (1 GPU): full GPU suite + smoke + model bench at HEAD
OUT=gpurun_out/${1:-r2final}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== model bench"; timeout 400 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-230
