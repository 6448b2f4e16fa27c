#!/bin/bash
OUT=gpurun_out/${1:-sanitize}; mkdir -p $OUT
cat > /tmp/san.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from allset_b200 import _lib
dev = torch.device('cuda:0')
d = 128
g = torch.Generator(device='cpu').manual_seed(0)
x = torch.randn(128 * 40 + 5, d, generator=g).to(dev)
w1 = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev); w2 = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
b = torch.randn(d, generator=g).to(dev)
ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
for ind, outd in ((torch.float32, torch.bfloat16), (torch.bfloat16, torch.float32)):
    _lib.mlp2_fwd(x.to(ind), w1, b, w2, b, ln, ln, True, outd)
    _lib.mlp2_fwd(x.to(ind), w1, b, None, None, ln, None, False, outd)
    _lib.pma_tail_fwd(x.to(ind), ln, w1, b, w2, b, ln, True, outd)
torch.cuda.synchronize()
print('done')
PY
echo "== memcheck"; timeout 400 compute-sanitizer --tool memcheck python /tmp/san.py 2>&1 | tail -8 | tee $OUT/memcheck.txt; compute-sanitizer --version | head -2
echo "== racecheck"; timeout 400 compute-sanitizer --tool racecheck python /tmp/san.py 2>&1 | tail -12 | tee $OUT/racecheck.txt
echo "== synccheck"; timeout 400 compute-sanitizer --tool synccheck python /tmp/san.py 2>&1 | tail -6 | tee $OUT/synccheck.txt
