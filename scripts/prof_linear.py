"""A few launches of each tcgen05 Linear kernel for an ncu capture (1 M rows, d = 128)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from allset_b200 import _lib
dev = torch.device('cuda:0')
rows, d = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 128
for dt in (torch.float32, torch.bfloat16):
    x = torch.randn(rows, d, device=dev).to(dt)
    dy = torch.randn(rows, d, device=dev).to(dt)
    w = torch.randn(d, d, device=dev) / d ** 0.5
    for _ in range(2):
        _lib.linear_fwd(x, w)
        _lib.linear_wgrad(dy, x)
torch.cuda.synchronize()
