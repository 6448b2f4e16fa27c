#!/bin/bash
OUT=gpurun_out/${1:-sanitize7}; mkdir -p $OUT
cat > /tmp/san7.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from allset_b200 import _lib
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
d = int(sys.argv[1])
rows = 128 * 450 + 5
x = torch.randn(rows, d, generator=g).to(dev)
w1 = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev); w2 = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
b = torch.randn(d, generator=g).to(dev)
ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
_lib.mlp2_fwd(x.bfloat16(), w1, b, w2, b, ln, ln, True, torch.bfloat16)
torch.cuda.synchronize()
print('done two-layer', d)
PY
for d in 64 128; do
echo "== synccheck fused two-layer mlp2_fwd d=$d"; timeout 200 compute-sanitizer --tool synccheck --print-limit 1 python /tmp/san7.py $d 2>&1 | grep -v "Host Frame\|in /" | head -9 | cut -c1-260 | tee -a $OUT/synccheck.txt
done
