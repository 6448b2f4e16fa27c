#!/bin/bash
# round 2, call X (1 GPU): final validation at HEAD -- full GPU suite, smoke, the bench line
OUT=gpurun_out/${1:-r2x}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-160; tail -2 $OUT/bench.err
