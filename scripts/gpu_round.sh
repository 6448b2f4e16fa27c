#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list + one full capture.  Outputs -> gpurun_out/
# usage: scripts/gpu_round.sh [tag]   (run under `gpurun --timeout 1800 -- bash scripts/gpu_round.sh r1a`)
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench" ; timeout 600 python bench.py --steps 50 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json ; tail -5 $OUT/bench.err
echo "== bench reference" ; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tee $OUT/bench_reference.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'segreduce|pma_|csr_|rowdot' -c 60 --csv \
  --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
tail -3 $OUT/ncu_launches.log
echo "== ncu full (segreduce)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segreduce_stream|segreduce_group' -s 2 -c 2 \
  -o $OUT/prof_segreduce -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-pma > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
