#!/bin/bash
# gpurun call: tcgen05 MLP check first (bounded), then the GPU suite, then the bench line.
OUT=gpurun_out/${1:-r1d}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== mlp2_check"; timeout 200 python scripts/mlp2_check.py --time 2>&1 | tail -40 | tee $OUT/mlp2_check.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json; tail -3 $OUT/bench.err
