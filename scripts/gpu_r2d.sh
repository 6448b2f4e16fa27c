#!/bin/bash
# round 2, call D (N GPUs, default 2): multi-GPU correctness + the exchange variants + the sharded SetGNN
N=${2:-2}
OUT=gpurun_out/${1:-r2d}; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== 2-GPU tests"
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 400 \
  -k "two_gpus" 2>&1 | tail -30 | tee $OUT/pytest_2gpu.txt
echo "== bench, bulk (TMA) peer stores"
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-mlp 2>$OUT/bench_bulk.err | tee $OUT/bench_bulk.json | cut -c1-200; tail -3 $OUT/bench_bulk.err
echo "== bench, direct peer stores"
ALLSET_PUSH=direct timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-mlp --no-e2e --no-pma 2>$OUT/bench_direct.err | tee $OUT/bench_direct.json | cut -c1-200; tail -3 $OUT/bench_direct.err
echo "== bench, bulk stores to the multicast address"
ALLSET_MULTICAST=1 timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-mlp --no-e2e --no-pma 2>$OUT/bench_bulk_mc.err | tee $OUT/bench_bulk_mc.json | cut -c1-200; tail -3 $OUT/bench_bulk_mc.err
echo "== bench, direct stores to the multicast address"
ALLSET_MULTICAST=1 ALLSET_PUSH=direct timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-mlp --no-e2e --no-pma 2>$OUT/bench_direct_mc.err | tee $OUT/bench_direct_mc.json | cut -c1-200; tail -3 $OUT/bench_direct_mc.err
echo "== sharded SetGNN"
timeout 400 $TR scripts/sharded_bench.py --config small --layers 2 --train --verify 2>$OUT/sharded_small.err | tee $OUT/sharded_small.json; tail -3 $OUT/sharded_small.err
timeout 600 $TR scripts/sharded_bench.py --config cfg4 --layers 1 --train 2>$OUT/sharded_cfg4.err | tee $OUT/sharded_cfg4.json; tail -3 $OUT/sharded_cfg4.err
ls $OUT
