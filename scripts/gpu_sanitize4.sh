#!/bin/bash
OUT=gpurun_out/${1:-sanitize4}; mkdir -p $OUT
cat > /tmp/san5.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from allset_b200 import _lib, ops
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
which, d = sys.argv[1], int(sys.argv[2])
rows = 128 * 450 + 5
x = torch.randn(rows, d, generator=g).to(dev)
w = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
b = torch.randn(d, generator=g).to(dev)
ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
if which == 'split_ln': _lib.linear_fwd(x, w, b, ln=ln, relu=True)
elif which == 'split_t': _lib.linear_fwd(x, w, transposed=True)
elif which == 'split': _lib.linear_fwd(x, w)
elif which == 'bf16': _lib.linear_fwd(x.bfloat16(), w, b, relu=True)
elif which == 'wgrad': _lib.linear_wgrad(x, x)
elif which == 'wgrad_bf16': _lib.linear_wgrad(x.bfloat16(), x.bfloat16())
elif which == 'rowop':
    xr = x.clone().requires_grad_(True)
    y = ops.rowop(xr, b, True, None, ln[0].clone().requires_grad_(True), ln[1], 1e-5, False, 0.5, torch.float32)
    y.sum().backward()
torch.cuda.synchronize()
print('done', which, d)
PY
for k in "split_ln 128" "split_t 128" "wgrad_bf16 128" "split 64" "bf16 64" "wgrad 64" "wgrad_bf16 64" "rowop 128" "rowop 64"; do
echo "== synccheck $k"; timeout 200 compute-sanitizer --tool synccheck --print-limit 2 python /tmp/san5.py $k 2>&1 | grep -v "Host Frame\|in /\|^=========         \|^=========     at " | head -14 | cut -c1-300 | tee -a $OUT/synccheck.txt
done
