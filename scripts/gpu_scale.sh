#!/bin/bash
# scaling runs on one box: scripts/gpu_scale.sh tag N [N...]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for N in "$@"; do
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps 30 --warmup 5 2>$OUT/n$N.err | tee $OUT/bench_n$N.json | cut -c1-300
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 30 --warmup 5 2>$OUT/n$N.err | tee $OUT/bench_n$N.json | cut -c1-300
  fi
  grep -vE "OMP_NUM|^\*+$|^$" $OUT/n$N.err | tail -4
done
