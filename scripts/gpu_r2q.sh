#!/bin/bash
# round 2, call Q (1 GPU): validation at HEAD + refreshed evidence (ncu pages exported on the box as CSV)
OUT=gpurun_out/${1:-r2q}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-160; tail -2 $OUT/bench.err
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee $OUT/bench_reference.json | cut -c1-200
echo "== kbench"; timeout 300 python scripts/kbench.py 2>&1 | grep '^{' | tee $OUT/kbench_default.json | cut -c1-300
timeout 300 python scripts/kbench.py 10000000 2000000 30 64 2>&1 | grep '^{' | tee $OUT/kbench_d64.json | cut -c1-300
KB_GRAPH=powerlaw timeout 300 python scripts/kbench.py 20000000 3200000 0 256 2>&1 | grep '^{' | tee $OUT/kbench_cfg5_scaled.json | cut -c1-300
echo "== linear bench"; timeout 300 python scripts/linear_bench.py 2>&1 | tee $OUT/linear_bench.jsonl | cut -c1-200
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'segreduce|pma_|mlp2_|csr_|rowdot|wgrad|fwd_kernel|bwd_kernel' -c 100 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -1 $OUT/ncu_launches.log | cut -c1-100
echo "== ncu full: tcgen05 Linear kernels"
T=/tmp/prof_linear
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp2_ws_kernel|wgrad_kernel' -c 8 -o $T -f python scripts/prof_linear.py > $OUT/ncu_linear.log 2>&1; tail -1 $OUT/ncu_linear.log
ncu -i $T.ncu-rep --page raw --csv > $OUT/linear_raw.csv 2>/dev/null
echo "== ncu full: one bf16-mode training step (rowop fwd / bwd, Linear kernels, aggregation fwd / bwd)"
T=/tmp/prof_train
timeout 600 ncu --set full --clock-control none -k regex:'fwd_kernel|bwd_kernel|mlp2_ws|wgrad_kernel|segreduce|pma_' -s 120 -c 60 -o $T -f \
  python scripts/prof_train.py 3 bf16only > $OUT/ncu_train.log 2>&1; tail -1 $OUT/ncu_train.log
ncu -i $T.ncu-rep --page raw --csv > $OUT/train_raw.csv 2>/dev/null
ls -la $OUT
