#!/bin/bash
# round 2, call F (1 GPU): full suite, ncu of the rowop kernels, kernel bench after the flush tightening
OUT=gpurun_out/${1:-r2f}; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 --tb=short > $OUT/pytest_gpu.txt 2>&1; tail -25 $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== kbench"; timeout 300 python scripts/kbench.py 2>&1 | tail -1 | tee $OUT/kbench_default.json
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
echo "== ncu rowop"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|bwd_kernel' -s 2 -c 2 -o $OUT/prof_rowop -f python /tmp/rowop_prof.py > $OUT/ncu_rowop.log 2>&1; tail -2 $OUT/ncu_rowop.log
echo "== training-step profile"; timeout 600 python scripts/prof_train.py 10 > $OUT/prof_train.txt 2>&1; grep "====" $OUT/prof_train.txt
ls -la $OUT
