"""Generate the golden vectors under tests/golden/ by running the REFERENCE ITSELF.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py            (dev container only: needs /root/reference)

The reference ships no tests, golden vectors or fixed seeds for this path, so parity is
pinned against outputs of the reference's own unmodified `src/layers.py`, `src/models.py`, `src/preprocessing.py`
and `src/load_other_datasets.py`, imported through `oracle/ref_harness.py` (third-party imports satisfied by
`oracle/shims/`).  Every fixture holds inputs, the reference `state_dict()`, outputs, per-half-layer intermediates
(forward hooks on the reference's V2EConvs / E2VConvs) and gradients, as plain tensors in a dict saved with
`torch.save` (loadable with `weights_only=True`).  `/root/reference` cannot travel to the GPU box; these files do.

Fixtures
  cora_alldeepsets.pt            BASELINE.json configs[0]: real cora cocitation through the reference preprocessing,
                                 AllDeepSets d=64 L=1 (src/run_one_model.sh:39-55 flags)
  citeseer_allsettransformer.pt  configs[1]: real citeseer cocitation, AllSetTransformer d=128 heads=4 L=2
  layers_small.pt                reference HalfNLHconv / PMA modules on small random bipartite graphs: sum / add /
                                 mean, float and int64 `norm`, heads 1/2/4/8, attention weights, interior empty
                                 segments, gradients
  setgnn_variants.pt             small SetGNN variants: GPR, LearnMask, mean, bn / None normalisation,
                                 MLP_num_layers 0/1/3, All_num_layers 0
"""
from __future__ import annotations

import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import ref_harness  # noqa: E402
from allset_oracle import config_namespace  # noqa: E402

OUT = os.path.join(os.path.dirname(_HERE), 'tests', 'golden')


def dense_to_sparse(x: torch.Tensor):
    """cora / citeseer features are > 98 % zeros: store coordinates + values, not 15-50 MB of zeros."""
    nz = x.nonzero(as_tuple=False)
    return {'shape': list(x.shape), 'rows': nz[:, 0].int(), 'cols': nz[:, 1].int(), 'vals': x[nz[:, 0], nz[:, 1]].clone()}


def run_setgnn(mods, args, x, edge_index, norm, seed, with_grads=True, tap_stride=1):
    """Build the reference SetGNN under a fixed seed, run eval forward (+ backward of a fixed scalar loss)."""
    torch.manual_seed(seed)
    model = mods.models.SetGNN(args, norm) if args.LearnMask else mods.models.SetGNN(args)
    model.reset_parameters()
    # make otherwise-trivial parameters non-trivial so the comparison exercises them
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith('normalizations.0.weight') or 'ln0.weight' in name or 'ln1.weight' in name:
                p.add_(0.1 * torch.randn_like(p))
            if name == 'Importance':
                p.mul_(0.5 + torch.rand_like(p))
    model.eval()
    taps = []
    hooks = []
    for i in range(len(model.V2EConvs)):
        for conv in (model.V2EConvs[i], model.E2VConvs[i]):
            hooks.append(conv.register_forward_hook(lambda m, inp, out: taps.append(out.detach().clone())))
    data = type('D', (), {})()
    data.x = x.clone().requires_grad_(with_grads)
    data.edge_index = edge_index.clone()
    data.norm = norm.clone()
    out = model(data)
    for h in hooks:
        h.remove()
    rec = {
        'state_dict': {k: v.detach().clone() for k, v in model.state_dict().items()},
        'logits': out.detach().clone(),
        # conv outputs BEFORE the outer relu of SetGNN.forward (models.py:475,478); stride-subsampled rows
        'taps': [t[::tap_stride].clone() for t in taps],
        'tap_stride': tap_stride,
        'edge_index_after': data.edge_index.clone(),      # the reference zero-bases row 1 in place (models.py:453-454)
    }
    if with_grads:
        torch.manual_seed(seed + 1)
        g = torch.randn_like(out)
        (out * g).sum().backward()
        rec['grad_logits'] = g
        rec['grad_x_rowsum'] = data.x.grad.sum(dim=1).clone()       # [N]: compact witness of d loss / d x
        rec['grads'] = {k: p.grad.detach().clone() for k, p in model.named_parameters()
                        if p.grad is not None and p.numel() <= 70000}
    return rec


def args_dict(args):
    return {k: v for k, v in vars(args).items()}


def real_dataset(mods, name, args_kw, seed, tap_stride):
    data = ref_harness.load_cocitation(name)
    n_classes = int(data.y.max()) + 1
    args = config_namespace(num_features=data.x.shape[1], num_classes=n_classes, **args_kw)
    rec = run_setgnn(mods, args, data.x, data.edge_index, data.norm, seed, tap_stride=tap_stride)
    rec.update({
        'name': name, 'args': args_dict(args), 'x_sparse': dense_to_sparse(data.x),
        'edge_index': data.edge_index.clone(), 'norm': data.norm.clone(), 'y': data.y.clone(),
        'n_nodes': int(data.x.shape[0]),
    })
    return rec


def random_bipartite(n_src, n_tgt, nnz, gen, empty_tgt=()):
    """COO [2, nnz] row 0 = source ids (ascending, like ExtractV2E output), row 1 = target ids; the last target id
    is always present (torch_scatter sizes its output by index.max()+1) and `empty_tgt` ids never are."""
    src = torch.randint(0, n_src, (nnz,), generator=gen)
    allowed = torch.tensor([t for t in range(n_tgt) if t not in set(empty_tgt)])
    tgt = allowed[torch.randint(0, allowed.numel(), (nnz,), generator=gen)]
    tgt[-1] = n_tgt - 1
    src[0] = n_src - 1
    order = torch.argsort(src, stable=True)
    return torch.stack([src[order], tgt[order]])


def layer_cases(mods):
    gen = torch.Generator().manual_seed(1234)
    cases = []

    def record(kind, module, x, ei, norm, aggr, extra):
        module.eval()
        x = x.clone().requires_grad_(True)
        rec = {'kind': kind, 'x': x.detach().clone(), 'edge_index': ei.clone(), 'extra': extra,
               'state_dict': {k: v.detach().clone() for k, v in module.state_dict().items()}}
        if kind == 'pma':
            out, (ei_back, alpha) = module(x, ei, return_attention_weights=True)
            rec['alpha'] = alpha.detach().clone()
        else:
            rec['norm'] = norm.clone()
            rec['aggr'] = aggr
            out = module(x, ei, norm, aggr)
        g = torch.randn(out.shape, generator=gen)
        (out * g).sum().backward()
        rec['out'] = out.detach().clone()
        rec['grad_out'] = g
        rec['grad_x'] = x.grad.detach().clone()
        rec['grads'] = {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}
        cases.append(rec)

    # HalfNLHconv, attention=False (AllDeepSets half layer, layers.py:623-656)
    for (n_src, n_tgt, nnz, d_in, d, L, aggr, norm_kind, normalization, empty) in [
        (40, 15, 120, 12, 16, 2, 'add', 'ones_i64', 'ln', ()),
        (40, 15, 120, 12, 16, 2, 'sum', 'float', 'ln', (3, 7)),
        (33, 21, 90, 10, 24, 2, 'mean', 'float', 'ln', (0, 5)),
        (33, 21, 90, 10, 20, 1, 'mean', 'ones_i64', 'None', ()),
        (25, 9, 60, 8, 8, 0, 'sum', 'float', 'ln', (2,)),
        (50, 12, 400, 6, 32, 3, 'add', 'float', 'bn', ()),
    ]:
        torch.manual_seed(len(cases))
        out_dim = d if L > 0 else d_in
        m = mods.layers.HalfNLHconv(in_dim=d_in, hid_dim=d, out_dim=out_dim, num_layers=L, dropout=0.3,
                                    Normalization=normalization, InputNorm=True, heads=1, attention=False)
        ei = random_bipartite(n_src, n_tgt, nnz, gen, empty)
        x = torch.randn(n_src, d_in, generator=gen)
        norm = torch.ones(nnz, dtype=torch.int64) if norm_kind == 'ones_i64' else torch.rand(nnz, generator=gen) + 0.5
        record('deepsets', m, x, ei, norm, aggr,
               {'in_dim': d_in, 'hid_dim': d, 'out_dim': out_dim, 'num_layers': L, 'Normalization': normalization,
                'InputNorm': True, 'n_tgt': n_tgt})

    # PMA (AllSetTransformer half layer, layers.py:42-199)
    for (n_src, n_tgt, nnz, d_in, d, heads, L, empty) in [
        (40, 15, 120, 12, 16, 1, 2, ()),
        (40, 15, 120, 12, 16, 4, 2, (3, 7)),
        (33, 21, 200, 10, 32, 8, 2, (0,)),
        (33, 21, 90, 10, 24, 2, 1, ()),
        (20, 6, 300, 16, 64, 4, 2, ()),
        (30, 10, 80, 9, 12, 4, 2, (1,)),      # C = 3: a 16-byte chunk would straddle heads
    ]:
        torch.manual_seed(100 + len(cases))
        m = mods.layers.PMA(d_in, d, d, L, heads=heads)
        with torch.no_grad():
            m.ln0.weight.add_(0.1 * torch.randn_like(m.ln0.weight))
            m.ln1.bias.add_(0.1 * torch.randn_like(m.ln1.bias))
        ei = random_bipartite(n_src, n_tgt, nnz, gen, empty)
        x = torch.randn(n_src, d_in, generator=gen) * 2.0
        record('pma', m, x, ei, None, None,
               {'in_channels': d_in, 'hid_dim': d, 'out_channels': d, 'num_layers': L, 'heads': heads, 'n_tgt': n_tgt})
    return cases


def variant_cases(mods):
    gen = torch.Generator().manual_seed(4321)
    out = []
    N, M, F, classes = 60, 25, 14, 5
    # reference layout: row 0 node ids ascending, row 1 hyperedge ids in [N, N+M) (preprocessing.py:394-447)
    ei = random_bipartite(N, M, 260, gen)
    ei[1] += N
    x = torch.randn(N, F, generator=gen)
    norm_ones = torch.ones(ei.shape[1], dtype=torch.int64)
    norm_f = torch.rand(ei.shape[1], generator=gen) + 0.5
    variants = [
        ('deepsets_mean_ln', dict(PMA=False, aggregate='mean', All_num_layers=2, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('deepsets_add_floatnorm', dict(PMA=False, aggregate='add', All_num_layers=1, MLP_hidden=16, Classifier_hidden=16), norm_f),
        ('deepsets_bn', dict(PMA=False, aggregate='add', normalization='bn', All_num_layers=1, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('deepsets_nonorm_mlp1', dict(PMA=False, aggregate='sum', normalization='None', MLP_num_layers=1, All_num_layers=2, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('deepsets_mlp0', dict(PMA=False, aggregate='mean', MLP_num_layers=0, All_num_layers=1, MLP_hidden=F, Classifier_hidden=16), norm_ones),
        ('deepsets_mlp3_noinputnorm', dict(PMA=False, aggregate='add', MLP_num_layers=3, deepset_input_norm=False, All_num_layers=1, MLP_hidden=16, Classifier_hidden=16), norm_f),
        ('deepsets_gpr', dict(PMA=False, aggregate='add', GPR=True, All_num_layers=2, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('deepsets_learnmask', dict(PMA=False, aggregate='mean', LearnMask=True, All_num_layers=1, MLP_hidden=16, Classifier_hidden=16), norm_f),
        ('transformer_h1', dict(PMA=True, heads=1, All_num_layers=1, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('transformer_h4_l2', dict(PMA=True, heads=4, All_num_layers=2, MLP_hidden=32, Classifier_hidden=16, Classifier_num_layers=1), norm_ones),
        ('transformer_gpr_h2', dict(PMA=True, heads=2, GPR=True, All_num_layers=2, MLP_hidden=16, Classifier_hidden=16), norm_ones),
        ('classifier_only', dict(PMA=True, All_num_layers=0, Classifier_hidden=16), norm_ones),
    ]
    for i, (name, kw, norm) in enumerate(variants):
        args = config_namespace(num_features=F, num_classes=classes, **kw)
        rec = run_setgnn(mods, args, x, ei, norm, seed=50 + i)
        rec.update({'name': name, 'args': args_dict(args), 'x': x.clone(), 'edge_index': ei.clone(), 'norm': norm.clone(),
                    'n_nodes': N})
        out.append(rec)
    return out


def preprocessing_cases(mods):
    """Inputs / outputs of the reference's own ExtractV2E, Add_Self_Loops, norm_contruction (preprocessing.py:394-464)."""
    import copy
    import io
    import tempfile
    import warnings
    import zipfile
    from types import SimpleNamespace
    cases = []

    def run(name, raw_ei, n_x, n_he):
        d = SimpleNamespace(edge_index=raw_ei.clone(), n_x=torch.tensor([n_x]), num_hyperedges=torch.tensor([n_he]))
        d = mods.preprocessing.ExtractV2E(d)
        v2e = d.edge_index.clone()
        d = mods.preprocessing.Add_Self_Loops(d)
        with_loops = d.edge_index.clone()
        tot = int(d.totedges)
        ones = mods.preprocessing.norm_contruction(copy.copy(d), option='all_one').norm.clone()
        sym = mods.preprocessing.norm_contruction(copy.copy(d), option='deg_half_sym').norm.clone()
        sym_noloop = mods.preprocessing.norm_contruction(SimpleNamespace(edge_index=v2e.clone()), option='deg_half_sym').norm.clone()
        cases.append({'name': name, 'raw': raw_ei.clone(), 'n_x': n_x, 'num_hyperedges': n_he, 'v2e': v2e,
                      'with_loops': with_loops, 'totedges': tot, 'norm_all_one': ones, 'norm_deg_half_sym': sym,
                      'norm_deg_half_sym_noloop': sym_noloop})

    # real datasets: the raw star expansion exactly as load_citation_dataset builds it
    for name in ('cora', 'citeseer'):
        with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ref_harness._RAW_ZIP) as z:
            base = 'AllSet_all_raw_data/cocitation/%s/' % name
            os.makedirs(os.path.join(tmp, name))
            for f in ('features.pickle', 'labels.pickle', 'hypergraph.pickle'):
                with open(os.path.join(tmp, name, f), 'wb') as out:
                    out.write(z.read(base + f))
            with warnings.catch_warnings(), ref_harness._quiet():
                warnings.simplefilter('ignore')
                data = mods.loaders.load_citation_dataset(path=tmp, dataset=name)
        run(name, data.edge_index, int(data.n_x), int(data.num_hyperedges))
    # random star expansions with singleton hyperedges, repeated members and nodes of degree 1
    gen = torch.Generator().manual_seed(99)
    for i, (n, m, nnz) in enumerate([(30, 12, 70), (200, 90, 500), (500, 400, 900)]):
        node = torch.randint(0, n, (nnz,), generator=gen)
        he = torch.randint(0, m, (nnz,), generator=gen)
        he[:m] = torch.arange(m)                       # every hyperedge id present (ids must be contiguous)
        node[:n] = torch.arange(n) if nnz >= n else node[:n]
        k = min(m, 8)
        # make the last k hyperedges singletons
        single = he >= m - k
        keep = ~single
        keep[torch.arange(nnz)[single][:0]] = True
        first = {}
        for j in range(nnz):
            if single[j]:
                e = int(he[j])
                if e not in first:
                    first[e] = j
                    keep[j] = True
        node, he = node[keep], he[keep]
        vv = torch.cat([node, he + n]); ee = torch.cat([he + n, node])
        raw = torch.stack([vv, ee])
        # coalesce like torch_sparse.coalesce (sorted by (row, col), duplicates removed)
        key = raw[0] * (n + m) + raw[1]
        key = torch.unique(key)
        raw = torch.stack([key // (n + m), key % (n + m)])
        run('random%d' % i, raw, n, m)
    return cases


def main():
    if not ref_harness.available():
        raise SystemExit('reference tree not found: golden vectors can only be generated in the dev container')
    torch.set_num_threads(1)               # sequential scatter_add_ order => reproducible bits
    os.makedirs(OUT, exist_ok=True)
    mods = ref_harness.load()

    rec = real_dataset(mods, 'cora', dict(PMA=False, aggregate='add', All_num_layers=1, MLP_num_layers=2, MLP_hidden=64,
                                          Classifier_num_layers=1, Classifier_hidden=64), seed=0, tap_stride=1)
    torch.save(rec, os.path.join(OUT, 'cora_alldeepsets.pt'))
    print('cora', rec['logits'].shape, [t.shape for t in rec['taps']])

    rec = real_dataset(mods, 'citeseer', dict(PMA=True, heads=4, All_num_layers=2, MLP_num_layers=2, MLP_hidden=128,
                                              Classifier_num_layers=1, Classifier_hidden=128), seed=0, tap_stride=8)
    torch.save(rec, os.path.join(OUT, 'citeseer_allsettransformer.pt'))
    print('citeseer', rec['logits'].shape, [t.shape for t in rec['taps']])

    cases = layer_cases(mods)
    torch.save(cases, os.path.join(OUT, 'layers_small.pt'))
    print('layers_small', len(cases))

    cases = variant_cases(mods)
    torch.save(cases, os.path.join(OUT, 'setgnn_variants.pt'))
    print('setgnn_variants', len(cases))
    cases = preprocessing_cases(mods)
    torch.save(cases, os.path.join(OUT, 'preprocessing.pt'))
    print('preprocessing', len(cases), [(c['name'], tuple(c['with_loops'].shape), c['totedges']) for c in cases])
    for f in sorted(os.listdir(OUT)):
        print('%-36s %8.2f MB' % (f, os.path.getsize(os.path.join(OUT, f)) / 1e6))


if __name__ == '__main__':
    main()
