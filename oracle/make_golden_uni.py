"""Fixtures for allset_b200.uni from the reference's OWN UniGCNII / UniGCNIIConv (reference src/models.py:909-995), run
unmodified under oracle/ref_harness (torch_scatter shim).  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_uni.py    ->  tests/golden/unigcnii.pt

Inputs follow train.py:390-418: (V, E) = COO of the incidence matrix, degV = vertex degree ^ -1/2 (inf -> 1),
degE = (mean vertex degree of the hyperedge) ^ -1/2."""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'unigcnii.pt')


def graph(n, m, max_size, seed):
    g = torch.Generator().manual_seed(seed)
    vs, es = [], []
    for e in range(m):
        s = int(torch.randint(1, max_size + 1, (1,), generator=g))
        vs.append(torch.randperm(n, generator=g)[:s])
        es.append(torch.full((s,), e, dtype=torch.long))
    V, E = torch.cat(vs), torch.cat(es)
    order = torch.sort(V, stable=True)[1]
    return V[order], E[order]


def case(mods, name, n, m, max_size, nfeat, nhid, nhead, nclass, nlayer, use_norm, seed):
    scatter = sys.modules['torch_scatter'].scatter
    V, E = graph(n, m, max_size, seed)
    degV = torch.bincount(V, minlength=n).view(-1, 1).float()
    degE = scatter(degV[V], E, dim=0, reduce='mean').pow(-0.5)
    degV = degV.pow(-0.5)
    degV[torch.isinf(degV)] = 1
    args = SimpleNamespace(UniGNN_degV=degV, UniGNN_degE=degE, UniGNN_use_norm=use_norm)
    torch.manual_seed(seed)
    model = mods.models.UniGCNII(args, nfeat=nfeat, nhid=nhid, nclass=nclass, nlayer=nlayer, nhead=nhead, V=V, E=E)
    model.reset_parameters()
    model.eval()
    x = torch.randn(n, nfeat, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    out = model(SimpleNamespace(x=x))
    gl = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 2))
    (out * gl).sum().backward()
    conv = model.convs[1]
    h = torch.randn(n, nhid * nhead, generator=torch.Generator().manual_seed(seed + 3))
    conv_out = conv(h, V, E, 0.1, 0.4, 0.5 * h)
    return {'name': name, 'V': V, 'E': E, 'degV': degV, 'degE': degE, 'use_norm': use_norm,
            'ctor': dict(nfeat=nfeat, nhid=nhid, nclass=nclass, nlayer=nlayer, nhead=nhead),
            'state_dict': {k: v.detach().clone() for k, v in model.state_dict().items()},
            'x': x.detach().clone(), 'logits': out.detach().clone(), 'grad_logits': gl, 'grad_x': x.grad.clone(),
            'grads': {k: p.grad.clone() for k, p in model.named_parameters()},
            'conv_in': h, 'conv_out': conv_out.detach().clone()}


def main():
    mods = ref_harness.load()
    cases = [case(mods, 'small', 300, 120, 9, 24, 16, 1, 5, 2, False, 7),
             case(mods, 'norm + heads', 500, 260, 14, 40, 8, 4, 7, 3, True, 8),
             case(mods, 'isolated nodes', 400, 60, 5, 16, 32, 1, 3, 1, False, 9)]
    torch.save(cases, OUT)
    for c in cases:
        print(c['name'], tuple(c['V'].shape), tuple(c['logits'].shape), list(c['state_dict'])[:3])
    print('%s %.2f MB' % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == '__main__':
    main()
