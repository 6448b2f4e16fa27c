#!/bin/bash
# round 2, call H (8 GPUs): the headline bench (default exchange + the bulk-store variant), configs[4] at full size,
# the sharded SetGNN at configs[3] / configs[4]
N=${2:-8}
OUT=gpurun_out/${1:-r2h}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== bench (default: direct stores, multicast X_e, selective X_v)"
timeout 300 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-mlp 2>$OUT/bench_default.err | tee $OUT/bench_default.json | cut -c1-160; tail -2 $OUT/bench_default.err
echo "== bench, bulk (TMA) stores"
ALLSET_PUSH=bulk timeout 200 $TR bench.py --gpus $N --steps 30 --warmup 5 --no-mlp --no-e2e --no-pma 2>$OUT/bench_bulk.err | tee $OUT/bench_bulk.json | cut -c1-160; tail -2 $OUT/bench_bulk.err
echo "== configs[4] raw op pair at full size: power-law |V|=50M |E|=8M d=256 bf16"
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --graph powerlaw --nodes 50000000 --hyperedges 8000000 --width 256 --no-mlp --no-e2e --no-pma 2>$OUT/bench_cfg5.err | tee $OUT/bench_cfg5.json | cut -c1-160; tail -2 $OUT/bench_cfg5.err
echo "== sharded SetGNN, configs[3] (AllSetTransformer heads=8, 10M / 2M)"
timeout 300 $TR scripts/sharded_bench.py --config cfg4 --layers 2 --train --steps 6 2>$OUT/sharded_cfg4.err | tee $OUT/sharded_cfg4.json; tail -2 $OUT/sharded_cfg4.err
echo "== sharded SetGNN, configs[4] (AllDeepSets d=256, 50M / 8M power law)"
timeout 400 $TR scripts/sharded_bench.py --config cfg5 --layers 1 --train --steps 4 2>$OUT/sharded_cfg5.err | tee $OUT/sharded_cfg5.json; tail -2 $OUT/sharded_cfg5.err
ls $OUT
