#!/bin/bash
# round 2, call C (1 GPU): the risky kernels FIRST under a short timeout (a hang must not burn the budget), then the full suite
OUT=gpurun_out/${1:-r2c}; mkdir -p $OUT
echo "== risky kernels first"
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -m gpu -x -q -p no:cacheprovider --timeout 120 \
  -k "cut_long or cut_high or fused_exchange or push_rows" 2>&1 | tail -30 | tee $OUT/pytest_risky.txt
rc=${PIPESTATUS[0]}
if [ "$rc" != "0" ]; then echo "RISKY TESTS FAILED rc=$rc -- stopping here"; nvidia-smi > $OUT/nvidia_smi_after.txt 2>&1; exit 1; fi
echo "== kbench powerlaw (cut kernels at scale)"
KB_GRAPH=powerlaw timeout 200 python scripts/kbench.py 20000000 3200000 0 256 2>&1 | tail -1 | tee $OUT/kbench_cfg5_scaled.json
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== training-step profile"; timeout 600 python scripts/prof_train.py 12 > $OUT/prof_train.txt 2>&1; grep "====" $OUT/prof_train.txt
echo "== model bench"; timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-200
echo "== kbench"
timeout 300 python scripts/kbench.py 10000000 2000000 30 64 2>&1 | tail -1 | tee $OUT/kbench_d64.json
timeout 300 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kbench_default.json
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300; tail -3 $OUT/bench.err
ls $OUT
