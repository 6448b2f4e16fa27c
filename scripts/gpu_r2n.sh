#!/bin/bash
# round 2, call N (1 GPU): full validation at HEAD + the ncu evidence of the round (pages exported on the box as CSV)
OUT=gpurun_out/${1:-r2n}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-160; tail -2 $OUT/bench.err
echo "== bench reference"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee $OUT/bench_reference.json | cut -c1-200
echo "== kbench"; timeout 300 python scripts/kbench.py 2>&1 | grep '^{' | tee $OUT/kbench_default.json | cut -c1-300
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'segreduce|pma_|mlp2_|csr_|rowdot|wgrad|rowop' -c 80 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
echo "== ncu full: stream kernels (sum + PMA), both directions"
T=/tmp/prof_stream
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segreduce_stream|pma_stream' -s 4 -c 4 -o $T -f \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-mlp > $OUT/ncu_stream.log 2>&1; tail -1 $OUT/ncu_stream.log
ncu -i $T.ncu-rep --page raw --csv > $OUT/stream_raw.csv 2>/dev/null
echo "== ncu full: training step kernels (rowop, tcgen05 Linear fwd / wgrad), fp32 mode first (AllDeepSets)"
T=/tmp/prof_train
timeout 600 ncu --set full --clock-control none -k regex:'rowop|mlp2_ws|wgrad_kernel' -s 50 -c 44 -o $T -f \
  python scripts/prof_train.py 3 deepsets > $OUT/ncu_train.log 2>&1; tail -1 $OUT/ncu_train.log
ncu -i $T.ncu-rep --page raw --csv > $OUT/train_raw.csv 2>/dev/null
ls -la $OUT
