#!/bin/bash
OUT=gpurun_out/${1:-sanitize3}; mkdir -p $OUT
cat > /tmp/san3.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from allset_b200 import _lib
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
which = sys.argv[1]
d = 128
rows = 128 * 450 + 5
x = torch.randn(rows, d, generator=g).to(dev)
w = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
if which == 'split':
    _lib.linear_fwd(x, w)
elif which == 'bf16':
    _lib.linear_fwd(x.bfloat16(), w)
elif which == 'wgrad':
    _lib.linear_wgrad(x, x)
torch.cuda.synchronize()
print('done', which)
PY
for k in bf16 split wgrad; do
echo "== synccheck $k"; timeout 300 compute-sanitizer --tool synccheck --print-limit 3 python /tmp/san3.py $k 2>&1 | grep -v "^=========     at\|Host Frame\|in /\|^=========         " | head -40 | tee $OUT/synccheck_$k.txt
done
cat > /tmp/san4.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import allset_b200
from allset_b200 import synthetic
dev = torch.device('cuda:0')
n, m = 400_000, 70_000
ei = synthetic.poisson_hypergraph(n, m, 8, seed=7, device=dev)
inc = allset_b200.Incidence.from_coo(ei[0], ei[1] - n, n_src=n, n_tgt=m)
xv = synthetic.features(n, 128, torch.bfloat16, seed=3, device=dev)
sc = torch.randn(n, 8, device=dev)
seed = torch.randn(1, 8, 16, device=dev)
out, _ = allset_b200.pma_aggregate(xv, sc, seed, inc, 8)
torch.cuda.synchronize()
print('done')
PY
echo "== racecheck pma"; timeout 500 compute-sanitizer --tool racecheck --print-limit 2 python /tmp/san4.py 2>&1 | cut -c1-260 | head -60 | tee $OUT/racecheck_pma.txt
