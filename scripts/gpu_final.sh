#!/bin/bash
# End-of-round validation: parity suite, smoke, bench (+ reference arm), model bench, tcgen05 check, ncu captures.
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-160; tail -2 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee $OUT/bench_reference.json | cut -c1-200
echo "== model bench"; timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-200
echo "== mlp2 check"; timeout 200 python scripts/mlp2_check.py --time 2>&1 | tee $OUT/mlp2_check.txt | grep timing | cut -c1-200
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'segreduce|pma_|mlp2_|csr_|rowdot' -c 80 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -2 $OUT/ncu_launches.log
echo "== ncu full (mlp2)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mlp2_ -s 2 -c 1 -o $OUT/prof_mlp2 -f python scripts/mlp2_check.py --profile > $OUT/ncu_mlp2.log 2>&1
tail -1 $OUT/ncu_mlp2.log
ls $OUT
