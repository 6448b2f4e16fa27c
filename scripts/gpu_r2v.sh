#!/bin/bash
# round 2, call V (8 GPUs): configs[4] raw op pair at full size on 8 ranks; the bench at N=4 and N=8 with the final code
OUT=gpurun_out/${1:-r2v}; mkdir -p $OUT
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552"
echo "== configs[4] raw op pair at full size: power-law |V|=50M |E|=8M d=256 bf16, N=8"
timeout 400 $TR8 bench.py --gpus 8 --steps 10 --warmup 3 --graph powerlaw --nodes 50000000 --hyperedges 8000000 --width 256 --no-mlp --no-e2e 2>$OUT/bench_cfg5.err | tee $OUT/bench_cfg5.json | cut -c1-200; tail -2 $OUT/bench_cfg5.err
echo "== bench N=4"
timeout 300 $TR4 bench.py --gpus 4 --steps 30 --warmup 5 --no-mlp 2>$OUT/bench_n4.err | tee $OUT/bench_n4.json | cut -c1-200; tail -2 $OUT/bench_n4.err
echo "== bench N=8"
timeout 300 $TR8 bench.py --gpus 8 --steps 30 --warmup 5 --no-mlp 2>$OUT/bench_n8.err | tee $OUT/bench_n8.json | cut -c1-200; tail -2 $OUT/bench_n8.err
ls $OUT
