"""Diagnostic: where do the fp32-mode gradients of the citeseer golden model differ from the reference's when the square
Linears run in split precision on tcgen05?  Compares x.grad row sums of (a) the cuBLAS path, (b) each tcgen05 kernel
switched on alone, (c) all of them, against the recorded fp32 reference AND an fp64 evaluation of the oracle (the
arbiter: two fp32 evaluations that differ by ReLU sign flips at |pre-activation| ~ 1e-6 are both 'right')."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import allset_oracle as O
from allset_b200 import ops
from conftest import load_golden
from test_gpu_parity import _build, dev

name = sys.argv[1] if len(sys.argv) > 1 else 'citeseer_allsettransformer.pt'
agg = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == 'bf16') else None
ops.FUSED_DENSE_MIN_ROWS = 0
rec = load_golden(name)
gl = rec['grad_logits']


def run(parts):
    ops.TC_LINEAR_PARTS = set(parts)
    model, data = _build(rec, agg_dtype=agg)
    data.x.requires_grad_(True)
    out = model(data)
    (out * gl.to(dev())).sum().backward()
    grads = {k: p.grad.float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    return out.detach().float().cpu(), data.x.grad.sum(dim=1).cpu(), grads


# fp64 arbiter on the CPU oracle
params = {k: v.double().requires_grad_(True) for k, v in rec['state_dict'].items() if v.is_floating_point()}
from conftest import golden_x
x64 = golden_x(rec).double().requires_grad_(True)
a = rec['args']
logits64, _ = O.setgnn(params, x64, rec['edge_index'], rec['norm'], PMA=a['PMA'], heads=a['heads'], aggregate=a['aggregate'])
(logits64 * gl.double()).sum().backward()
t_rows = x64.grad.sum(dim=1)
ref_rows = rec['grad_x_rowsum'].double() if 'grad_x_rowsum' in rec else None


def bad(a_, b_, rtol=1e-3, atol=1e-4):
    a_, b_ = a_.double(), b_.double()
    return int(((a_ - b_).abs() > atol + rtol * b_.abs()).sum())


print('rows', t_rows.numel(), 'recorded fp32 reference vs fp64:', None if ref_rows is None else bad(ref_rows, t_rows))
for parts in ([], ['fwd'], ['dgrad'], ['wgrad'], ['fwd', 'dgrad', 'wgrad']):
    out, rows, grads = run(parts)
    line = 'parts=%-22s logits err vs fp64 %.2e | x.grad rows off vs fp64: %d' % (
        ','.join(parts) or 'cuBLAS', (out.double() - logits64.detach()).abs().max().item(), bad(rows, t_rows))
    if ref_rows is not None:
        line += ', vs recorded fp32: %d' % bad(rows, ref_rows)
    worst = max(((grads[k].double() - params[k].grad).norm().item() / (params[k].grad.norm().item() + 1e-30), k)
                for k in grads if params[k].grad is not None)
    line += ' | worst param grad rel-L2 vs fp64: %.2e (%s)' % worst
    print(line, flush=True)
