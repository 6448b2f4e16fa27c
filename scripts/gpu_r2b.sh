#!/bin/bash
# round 2, call B (1 GPU): full parity suite (new: rowop / training chain / long-segment cuts / push modes / sharded world 1),
# training-step profile, model bench, kernel bench on the power-law and d=64 shapes, bench
OUT=gpurun_out/${1:-r2b}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== training-step profile"; timeout 600 python scripts/prof_train.py 16 > $OUT/prof_train.txt 2>&1; grep "====" $OUT/prof_train.txt
echo "== model bench"; timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-260
echo "== kbench"
KB_GRAPH=powerlaw timeout 300 python scripts/kbench.py 20000000 3200000 0 256 2>&1 | tail -1 | tee $OUT/kbench_cfg5_scaled.json
timeout 300 python scripts/kbench.py 10000000 2000000 30 64 2>&1 | tail -1 | tee $OUT/kbench_d64.json
timeout 300 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kbench_default.json
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300; tail -3 $OUT/bench.err
ls $OUT
