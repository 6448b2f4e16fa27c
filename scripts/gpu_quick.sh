#!/bin/bash
OUT=gpurun_out/${1:-quick}; mkdir -p $OUT
echo "== mlp2_check"; timeout 200 python scripts/mlp2_check.py --time 2>&1 | tail -8 | cut -c1-330 | tee $OUT/mlp2_check.txt
echo "== pytest mlp2"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mlp2 or tcgen05" 2>&1 | tail -5 | tee $OUT/pytest_mlp2.txt
