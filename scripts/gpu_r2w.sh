#!/bin/bash
# round 2, call W (2 GPUs): sharded vs unsharded check on a power-law graph (long segments cut at partition-dependent boundaries)
OUT=gpurun_out/${1:-r2w}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --graph powerlaw --nodes 6000000 --hyperedges 1000000 --width 256 --no-mlp --no-e2e 2>$OUT/bench_pl.err | tee $OUT/bench_pl.json | cut -c1-300; grep -v "^\*\|^$\|NCCL\|OMP" $OUT/bench_pl.err | tail -12
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --graph powerlaw --nodes 6000000 --hyperedges 1000000 --width 256 --no-mlp --no-e2e --dtype f32 --no-pma 2>$OUT/bench_pl_f32.err | tee $OUT/bench_pl_f32.json | cut -c1-300; grep -v "^\*\|^$\|NCCL\|OMP" $OUT/bench_pl_f32.err | tail -12
