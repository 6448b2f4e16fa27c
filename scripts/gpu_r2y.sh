#!/bin/bash
# round 2, call Y (8 GPUs): configs[4] raw op pair at full size on 8 ranks
OUT=gpurun_out/${1:-r2y}; mkdir -p $OUT
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
timeout 300 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --graph powerlaw --nodes 50000000 --hyperedges 8000000 --width 256 --no-mlp --no-e2e 2>$OUT/bench_cfg5.err | tee $OUT/bench_cfg5.json | cut -c1-300; grep -v "^\*\|^$\|NCCL\|OMP" $OUT/bench_cfg5.err | tail -6
