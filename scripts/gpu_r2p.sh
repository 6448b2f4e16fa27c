#!/bin/bash
# round 2, call P (2 GPUs): the multi-GPU tests with two devices visible, the N=2 bench, the sharded SetGNN check, model bench
OUT=gpurun_out/${1:-r2p}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short \
  -k "multi or two_gpu or sharded or world or exchange or push" > $OUT/pytest_multi.txt 2>&1; tail -6 $OUT/pytest_multi.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-mlp 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-200; tail -2 $OUT/bench_n2.err
timeout 300 $TR scripts/sharded_bench.py --config small --layers 2 --train --verify --steps 6 2>$OUT/sharded_small.err | tee $OUT/sharded_small.json; tail -2 $OUT/sharded_small.err
timeout 300 $TR scripts/sharded_bench.py --config small --layers 2 --train --verify --steps 6 --dtype f32 2>$OUT/sharded_small_f32.err | tee $OUT/sharded_small_f32.json; tail -2 $OUT/sharded_small_f32.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python scripts/model_bench.py 2>&1 | grep '^{' | tee $OUT/model_bench.jsonl | cut -c1-250
