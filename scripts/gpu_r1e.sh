#!/bin/bash
OUT=gpurun_out/${1:-r1e}; mkdir -p $OUT
echo "== mlp2_check v2"; timeout 200 python scripts/mlp2_check.py --time 2>&1 | tail -40 | tee $OUT/mlp2_check_v2.txt
echo "== mlp2_check v1"; ALLSET_MLP2_V1=1 timeout 200 python scripts/mlp2_check.py --time 2>&1 | grep timing | tee $OUT/mlp2_check_v1.txt
echo "== pytest mlp2"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mlp2 or tcgen05" 2>&1 | tail -15 | tee $OUT/pytest_mlp2.txt
echo "== model bench"; timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl
echo "== ncu mlp2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp2_ -s 2 -c 1 -o $OUT/prof_mlp2 -f python scripts/mlp2_check.py --profile > $OUT/ncu_mlp2.log 2>&1
tail -2 $OUT/ncu_mlp2.log
