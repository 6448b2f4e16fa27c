#!/bin/bash
# round 2: compute-sanitizer on the kernels added this round (tcgen05 Linear forward split / bf16 / transposed, weight gradient,
# rowop forward / backward, PMA and sum stream kernels on a graph with cut long segments)
OUT=gpurun_out/${1:-sanitize2}; mkdir -p $OUT
cat > /tmp/san2.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import allset_b200
from allset_b200 import _lib, ops, synthetic
dev = torch.device('cuda:0')
g = torch.Generator(device='cpu').manual_seed(0)
for d in (128, 64):
    rows = 128 * 3 * 148 // 37 + 5                    # several tiles per CTA on some CTAs, a ragged last tile
    rows = 128 * 450 + 5
    x = torch.randn(rows, d, generator=g).to(dev)
    dy = torch.randn(rows, d, generator=g).to(dev)
    w = (torch.randn(d, d, generator=g) / d ** 0.5).to(dev)
    b = torch.randn(d, generator=g).to(dev)
    ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
    _lib.linear_fwd(x, w, b, ln=ln, relu=True)                      # split precision, LayerNorm prologue
    _lib.linear_fwd(dy, w, transposed=True)                         # split precision, input gradient
    _lib.linear_fwd(x.bfloat16(), w, b, relu=True)                  # bf16 operands
    _lib.linear_wgrad(dy, x)                                        # split precision
    _lib.linear_wgrad(dy.bfloat16(), x.bfloat16())
    xr = x.clone().requires_grad_(True)
    y = ops.rowop(xr, b, True, None, ln[0].clone().requires_grad_(True), ln[1], 1e-5, False, 0.5, torch.float32)
    y.sum().backward()
n, m = 400_000, 70_000
ei = synthetic.powerlaw_hypergraph(n, m, 2, 4096, 2.0, seed=7, device=dev)
inc = allset_b200.Incidence.from_coo(ei[0], ei[1] - n, n_src=n, n_tgt=m)
xv = synthetic.features(n, 128, torch.bfloat16, seed=3, device=dev)
xe = allset_b200.segment_reduce(xv, inc, None, 'sum')
sc = torch.randn(n, 8, device=dev)
seed = torch.randn(1, 8, 16, device=dev)
out, _ = allset_b200.pma_aggregate(xv, sc, seed, inc, 8)
e2v = inc.reversed()
out2, _ = allset_b200.pma_aggregate(xe, torch.randn(m, 8, device=dev), seed, e2v, 8)
torch.cuda.synchronize()
print('done')
PY
echo "== memcheck"; timeout 500 compute-sanitizer --tool memcheck python /tmp/san2.py 2>&1 | tail -6 | tee $OUT/memcheck.txt
echo "== racecheck"; timeout 700 compute-sanitizer --tool racecheck python /tmp/san2.py 2>&1 | tail -8 | tee $OUT/racecheck.txt
echo "== synccheck"; timeout 500 compute-sanitizer --tool synccheck python /tmp/san2.py 2>&1 | tail -6 | tee $OUT/synccheck.txt
