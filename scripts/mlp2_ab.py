"""A/B timing of the tcgen05 dense kernels at the bench size (10 M rows) and at 1 M rows; run once per environment setting."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from allset_b200 import _lib
dev = torch.device('cuda:0')


def t(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


d = 128
w1 = torch.randn(d, d, device=dev) / d ** 0.5
w2 = torch.randn(d, d, device=dev) / d ** 0.5
bz = torch.zeros(d, device=dev)
ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
res = {'env': {k: v for k, v in os.environ.items() if k.startswith('ALLSET_')}}
for rows in (10_000_000, 1_000_000):
    xb = torch.randn(rows, d, device=dev).bfloat16()
    res['mlp2_bf16_bf16_%dM' % (rows // 1_000_000)] = t(lambda: _lib.mlp2_fwd(xb, w1, bz, w2, bz, ln, ln, True, torch.bfloat16))
    res['linear_bf16_%dM' % (rows // 1_000_000)] = t(lambda: _lib.linear_fwd(xb, w1))
    if rows <= 4_000_000:
        xf = xb.float()
        res['mlp2_f32_f32_%dM' % (rows // 1_000_000)] = t(lambda: _lib.mlp2_fwd(xf, w1, bz, w2, bz, ln, ln, True, torch.float32))
        res['linear_split_%dM' % (rows // 1_000_000)] = t(lambda: _lib.linear_fwd(xf, w1))
        res['wgrad_split_%dM' % (rows // 1_000_000)] = t(lambda: _lib.linear_wgrad(xf, xf))
        res['wgrad_bf16_%dM' % (rows // 1_000_000)] = t(lambda: _lib.linear_wgrad(xb, xb))
        del xf
    del xb
print(json.dumps(res), flush=True)
