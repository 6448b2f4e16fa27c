"""Fixtures for allset_b200.baselines from the reference's OWN HCHA / HypergraphConv / HNHN / UniGNN classes (reference
src/layers.py:233-494, src/models.py:207-292,601-907), run unmodified under oracle/ref_harness (third-party shims).
TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_baselines.py    ->  tests/golden/baselines.pt
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'baselines.pt')


def graph(n, m, max_size, seed):
    """edge_index [2, nnz]: row 0 node ids ascending (stable), row 1 zero-based hyperedge ids; every hyperedge non-empty."""
    g = torch.Generator().manual_seed(seed)
    vs, es = [], []
    for e in range(m):
        s = int(torch.randint(1, max_size + 1, (1,), generator=g))
        vs.append(torch.randperm(n, generator=g)[:s])
        es.append(torch.full((s,), e, dtype=torch.long))
    V, E = torch.cat(vs), torch.cat(es)
    order = torch.sort(V, stable=True)[1]
    return torch.stack([V[order], E[order]])


def run(model, call, seed):
    out = call()
    gl = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 2))
    (out * gl).sum().backward()
    return {'state_dict': {k: v.detach().clone() for k, v in model.state_dict().items()},
            'out': out.detach().clone(), 'grad_out': gl,
            'grads': {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}}


def hcha_case(mods, name, n, m, F_in, hid, classes, layers, sym, seed):
    ei = graph(n, m, 9, seed)
    args = SimpleNamespace(All_num_layers=layers, dropout=0.5, HCHA_symdegnorm=sym, num_features=F_in, MLP_hidden=hid,
                           num_classes=classes)
    torch.manual_seed(seed)
    model = mods.models.HCHA(args)
    model.reset_parameters()
    model.eval()
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    rec = run(model, lambda: model(SimpleNamespace(x=x, edge_index=ei)), seed)
    rec.update(kind='HCHA', name=name, args=vars(args), x=x.detach().clone(), edge_index=ei, grad_x=x.grad.clone())
    return rec


def hconv_attention_case(mods, name, n, m, F_in, out_c, heads, concat, sym, seed):
    ei = graph(n, m, 7, seed)                       # the reference's attention path indexes node rows by hyperedge id: m <= n
    torch.manual_seed(seed)
    conv = mods.layers.HypergraphConv(F_in, out_c, symdegnorm=sym, use_attention=True, heads=heads, concat=concat)
    conv.eval()
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    rec = run(conv, lambda: conv(x, ei), seed)
    rec.update(kind='HypergraphConvAttention', name=name, ctor=dict(in_channels=F_in, out_channels=out_c, symdegnorm=sym,
               use_attention=True, heads=heads, concat=concat), x=x.detach().clone(), edge_index=ei, grad_x=x.grad.clone())
    return rec


def hnhn_case(mods, name, n, m, F_in, hid, classes, layers, nonlinear, seed):
    ei = graph(n, m, 8, seed)
    g = torch.Generator().manual_seed(seed + 5)
    args = SimpleNamespace(All_num_layers=layers, dropout=0.5, num_features=F_in, MLP_hidden=hid, num_classes=classes,
                           HNHN_nonlinear_inbetween=nonlinear)
    torch.manual_seed(seed)
    model = mods.models.HNHN(args)
    model.reset_parameters()
    model.eval()
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    extra = dict(D_v_beta=torch.rand(n, generator=g) + 0.2, D_e_beta_inv=torch.rand(m, generator=g) + 0.2,
                 D_e_alpha=torch.rand(m, generator=g) + 0.2, D_v_alpha_inv=torch.rand(n, generator=g) + 0.2)
    rec = run(model, lambda: model(SimpleNamespace(x=x, edge_index=ei, **extra)), seed)
    rec.update(kind='HNHN', name=name, args=vars(args), x=x.detach().clone(), edge_index=ei, extra=extra,
               grad_x=x.grad.clone())
    return rec


def unignn_case(mods, name, model_name, n, m, F_in, hid, heads, classes, layers, first, second, use_norm, seed):
    scatter = sys.modules['torch_scatter'].scatter
    ei = graph(n, m, 10, seed)
    V, E = ei[0].clone(), ei[1].clone()
    degV = torch.bincount(V, minlength=n).view(-1, 1).float()
    degE = scatter(degV[V], E, dim=0, reduce='mean').pow(-0.5)
    degV = degV.pow(-0.5)
    degV[torch.isinf(degV)] = 1
    args = SimpleNamespace(model_name=model_name, first_aggregate=first, second_aggregate=second, use_norm=use_norm,
                           degE=degE, degV=degV, attn_drop=0.0, input_drop=0.0, dropout=0.0, activation='relu')
    torch.manual_seed(seed)
    model = mods.models.UniGNN(args, F_in, hid, classes, layers, heads, V, E)
    for p in model.parameters():                      # UniGINConv.eps / att_* keep ctor values; give eps a non-trivial one
        if p.numel() == 1:
            p.data.fill_(0.3)
    model.eval()
    x = torch.randn(n, F_in, generator=torch.Generator().manual_seed(seed + 1), requires_grad=True)
    rec = run(model, lambda: model(x), seed)
    rec.update(kind='UniGNN', name=name, V=V, E=E, degV=degV, degE=degE, x=x.detach().clone(), grad_x=x.grad.clone(),
               args={k: v for k, v in vars(args).items() if k not in ('degE', 'degV')},
               ctor=dict(nfeat=F_in, nhid=hid, nclass=classes, nlayer=layers, nhead=heads))
    return rec


def main():
    mods = ref_harness.load()
    cases = [
        hcha_case(mods, 'hcha_l2', 300, 120, 24, 16, 5, 2, False, 11),
        hcha_case(mods, 'hgnn_symdegnorm_l3', 260, 90, 20, 32, 4, 3, True, 12),
        hcha_case(mods, 'hcha_l1_isolated_nodes', 400, 50, 16, 8, 3, 1, False, 13),
        hconv_attention_case(mods, 'hconv_att_concat', 200, 80, 12, 8, 4, True, False, 14),
        hconv_attention_case(mods, 'hconv_att_mean', 220, 70, 10, 6, 2, False, False, 15),
        hnhn_case(mods, 'hnhn_l2', 300, 110, 24, 16, 5, 2, True, 16),
        hnhn_case(mods, 'hnhn_l1_linear', 240, 60, 12, 8, 3, 1, False, 17),
        hnhn_case(mods, 'hnhn_l3', 200, 90, 10, 12, 4, 3, True, 18),
        unignn_case(mods, 'unisage_mean_sum', 'UniSAGE', 300, 120, 24, 8, 4, 5, 2, 'mean', 'sum', False, 21),
        unignn_case(mods, 'unisage_max_mean_norm', 'UniSAGE', 260, 100, 16, 8, 2, 4, 3, 'max', 'mean', True, 22),
        unignn_case(mods, 'unigin', 'UniGIN', 280, 90, 20, 8, 4, 6, 2, 'mean', 'sum', False, 23),
        unignn_case(mods, 'unigcn_norm', 'UniGCN', 300, 130, 18, 8, 2, 5, 3, 'mean', 'sum', True, 24),
        unignn_case(mods, 'unigcn2', 'UniGCN2', 240, 80, 14, 8, 2, 3, 2, 'sum', 'sum', False, 25),
        unignn_case(mods, 'unigat', 'UniGAT', 300, 120, 24, 8, 4, 5, 2, 'mean', 'sum', False, 26),
        unignn_case(mods, 'unigat_l3_norm', 'UniGAT', 220, 140, 12, 4, 8, 4, 3, 'sum', 'sum', True, 27),
    ]
    torch.save(cases, OUT)
    for c in cases:
        print(c['kind'], c['name'], tuple(c['out'].shape), len(c['state_dict']))
    print('%s %.2f MB' % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == '__main__':
    main()
