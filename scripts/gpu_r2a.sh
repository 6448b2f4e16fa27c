#!/bin/bash
# round 2, call A: parity tests, the reference's train.py through the drop-in, training-step profile, bench
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
CORA="--method AllDeepSets --dname cora --All_num_layers 1 --MLP_num_layers 2 --Classifier_num_layers 1 --MLP_hidden 64 --Classifier_hidden 64 --wd 0 --feature_noise 0.0 --lr 0.001 --epochs 500"
CITE="--method AllSetTransformer --dname citeseer --All_num_layers 2 --MLP_num_layers 2 --Classifier_num_layers 1 --MLP_hidden 128 --Classifier_hidden 128 --heads 4 --wd 0 --feature_noise 0.0 --lr 0.001 --epochs 500"
echo "== train.py through the drop-in"
timeout 600 python scripts/run_train.py --impl dropin -- $CORA --runs 10 --cuda 0 2>$OUT/train_dropin_cora.err | grep '^{' | tee $OUT/train_dropin_cora.json
timeout 600 python scripts/run_train.py --impl dropin --agg-dtype bf16 -- $CORA --runs 10 --cuda 0 2>$OUT/train_dropin_cora_bf16.err | grep '^{' | tee $OUT/train_dropin_cora_bf16.json
timeout 600 python scripts/run_train.py --impl dropin -- $CITE --runs 10 --cuda 0 2>$OUT/train_dropin_citeseer.err | grep '^{' | tee $OUT/train_dropin_citeseer.json
echo "== train.py, reference modules on the same GPU (ATen scatter_add_ under the shims)"
timeout 600 python scripts/run_train.py --impl reference -- $CORA --runs 5 --cuda 0 2>$OUT/train_ref_gpu_cora.err | grep '^{' | tee $OUT/train_ref_gpu_cora.json
timeout 600 python scripts/run_train.py --impl reference -- $CITE --runs 5 --cuda 0 2>$OUT/train_ref_gpu_citeseer.err | grep '^{' | tee $OUT/train_ref_gpu_citeseer.json
echo "== training-step profile at config-3 size"
timeout 600 python scripts/prof_train.py > $OUT/prof_train.txt 2>&1 ; tail -5 $OUT/prof_train.txt
echo "== model bench" ; timeout 600 python scripts/model_bench.py 2>&1 | grep "^{" | tee $OUT/model_bench.jsonl
echo "== bench" ; timeout 600 python bench.py --steps 30 --warmup 5 2>$OUT/bench.err | tee $OUT/bench.json ; tail -3 $OUT/bench.err
ls -la $OUT
