#!/bin/bash
# round 2, call E (1 GPU): failing tests with full output, bf16 error budget, ncu of the rowop kernels
OUT=gpurun_out/${1:-r2e}; mkdir -p $OUT
echo "== train-path / dropin tests"
timeout 900 python -m pytest tests/test_train_path.py tests/test_train_dropin.py tests/test_baselines.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 300 --tb=short \
  -k "train or baselines or forward_and_gradients or c_abi or linear_nb" > $OUT/pytest_train.txt 2>&1; tail -60 $OUT/pytest_train.txt
echo "== bf16 parity report"; timeout 300 python scripts/bf16_parity_report.py 2>&1 | grep '^{' | tee $OUT/bf16_parity.jsonl
echo "== ncu rowop"
cat > /tmp/rowop_prof.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from allset_b200 import _lib
n, d = 1_000_000, 128
x = torch.randn(n, d, device='cuda').bfloat16(); dy = torch.randn(n, d, device='cuda').bfloat16()
b = torch.randn(d, device='cuda'); g = torch.rand(d, device='cuda') + 0.5; be = torch.randn(d, device='cuda')
for _ in range(3):
    out, st = _lib.rowop_fwd(x, b, True, None, g, be, 1e-5, False, 0.5, 7, torch.bfloat16, want_stats=True)
    _lib.rowop_bwd(dy, x, b, True, None, g, be, st, False, 0.5, 7, False)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowop -s 4 -c 2 -o $OUT/prof_rowop -f python /tmp/rowop_prof.py > $OUT/ncu_rowop.log 2>&1; tail -2 $OUT/ncu_rowop.log
ls -la $OUT
