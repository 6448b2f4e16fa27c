#!/bin/bash
OUT=gpurun_out/${1:-n2gpu}; mkdir -p $OUT
echo "== pytest 2-GPU tests"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "two_gpus or fused_exchange" 2>&1 | tail -5 | tee $OUT/pytest_2gpu.txt
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-300; tail -3 $OUT/bench_n2.err
