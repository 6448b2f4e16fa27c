#!/bin/bash
# round 2, very last call (1 GPU): full GPU suite + smoke on the final build
OUT=gpurun_out/${1:-r2last}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
