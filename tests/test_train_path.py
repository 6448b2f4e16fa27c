"""Training path of the dense glue: rowop kernels (allset_rowop_fwd / _bwd: bias, ReLU, residual, LayerNorm, ReLU,
dropout in one pass each way, fp32 or bf16 rows) + bias-free tensor-core Linears, against plain torch fp32 autograd and
against the reference's recorded gradients (reference src/layers.py:571-579, 153-157; src/train.py:478-482)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from test_gpu_parity import _build, ab, assert_grad_close, dev

pytestmark = pytest.mark.gpu

VARIANTS = {
    'bias_relu_ln': dict(bias=True, relu=True, ln=True),
    'ln_only': dict(ln=True),
    'bias_only': dict(bias=True),
    'residual_ln_relu': dict(bias=True, relu=True, residual=True, ln=True, relu_out=True),
    'bias_relu': dict(bias=True, relu=True),
}


def _ref(x, bias, relu, residual, gamma, beta, relu_out, mask_scale):
    t = x if bias is None else x + bias
    if relu:
        t = F.relu(t)
    if residual is not None:
        t = t + residual
    if gamma is not None:
        t = F.layer_norm(t, (t.shape[1],), gamma, beta, 1e-5)
    if relu_out:
        t = F.relu(t)
    return t if mask_scale is None else t * mask_scale


@pytest.mark.parametrize('d', [64, 128, 256, 512, 1024])
@pytest.mark.parametrize('variant', sorted(VARIANTS))
@pytest.mark.parametrize('io', ['f32', 'bf16', 'f32_to_bf16'])
def test_rowop_forward_backward_vs_torch(d, variant, io):
    from allset_b200 import ops
    v = VARIANTS[variant]
    g = torch.Generator().manual_seed(d + len(variant))
    rows = 3001                                      # not a multiple of the CTA row count: grid-stride tail
    in_dt = torch.bfloat16 if io == 'bf16' else torch.float32
    out_dt = torch.float32 if io == 'f32' else torch.bfloat16
    x = (torch.randn(rows, d, generator=g) * 2).to(in_dt).to(dev())
    r = torch.randn(rows, d, generator=g).to(in_dt).to(dev()) if v.get('residual') else None
    b = torch.randn(d, generator=g).to(dev()) if v.get('bias') else None
    gam = (torch.rand(d, generator=g) + 0.5).to(dev()) if v.get('ln') else None
    bet = torch.randn(d, generator=g).to(dev()) if v.get('ln') else None
    dy = torch.randn(rows, d, generator=g).to(out_dt).to(dev())

    def leaves(ts, dtype=None):
        return [None if t is None else (t.clone() if dtype is None else t.to(dtype)).requires_grad_(True) for t in ts]

    xa, ra, ba, ga, bea = leaves([x, r, b, gam, bet])
    xb, rb, bb, gb, beb = leaves([x, r, b, gam, bet], torch.float32)
    out = ops.rowop(xa, ba, v.get('relu', False), ra, ga, bea, 1e-5, v.get('relu_out', False), 0.0, out_dt)
    ref = _ref(xb, bb, v.get('relu', False), rb, gb, beb, v.get('relu_out', False), None)
    assert out.dtype == out_dt
    lo = out_dt == torch.bfloat16
    torch.testing.assert_close(out.float(), ref, rtol=1e-2 if lo else 1e-5, atol=2e-2 if lo else 1e-5)
    (out.float() * dy.float()).sum().backward()
    (ref * dy.float()).sum().backward()
    # gradients: dy is exact in both; dx is rounded to the gradient dtype once
    glo = lo or in_dt == torch.bfloat16
    scale = xb.grad.abs().max().item()
    assert (xa.grad.float() - xb.grad).abs().max().item() <= (1e-2 if glo else 1e-4) * scale + 1e-5
    if r is not None:
        assert (ra.grad.float() - rb.grad).abs().max().item() <= (1e-2 if glo else 1e-4) * rb.grad.abs().max().item() + 1e-5
    for name, a_, b_ in (('bias', ba, bb), ('gamma', ga, gb), ('beta', bea, beb)):
        if a_ is not None:
            assert_grad_close(a_.grad, b_.grad, name, rel=1e-4)      # parameter sums are fp32 from exact inputs


@pytest.mark.parametrize('d', [64, 128, 512])
@pytest.mark.parametrize('p', [0.2, 0.5])
def test_rowop_dropout_is_bernoulli_and_regenerated_by_the_backward(d, p):
    from allset_b200 import ops
    g = torch.Generator().manual_seed(7 * d)
    rows = 20000
    x = (torch.randn(rows, d, generator=g) + 3.0).to(dev())            # strictly away from 0 after the LayerNorm shift
    gam = (torch.rand(d, generator=g) + 0.5).to(dev())
    bet = (torch.rand(d, generator=g) + 5.0).to(dev())                 # LN output > 0 everywhere: zeros are dropout only
    torch.manual_seed(11)
    xa = x.clone().requires_grad_(True)
    out = ops.rowop(xa, None, False, None, gam, bet, 1e-5, False, p, torch.float32)
    keep = out != 0
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 4e-3, frac
    # per-column and per-row keep rates are flat (no structure in the counter hash)
    assert (keep.float().mean(dim=0) - (1 - p)).abs().max().item() < 0.03
    assert (keep.float().mean(dim=1) - (1 - p)).abs().max().item() < 0.25 * (128 / d) ** 0.5
    ref_full = F.layer_norm(x, (d,), gam, bet, 1e-5)
    scale = 65536.0 / (65536 - round(p * 65536))
    torch.testing.assert_close(out[keep], (ref_full * scale)[keep], rtol=1e-5, atol=1e-5)
    # a second call draws a different mask; the same torch seed reproduces the first
    out2 = ops.rowop(x, None, False, None, gam, bet, 1e-5, False, p, torch.float32)
    assert not torch.equal(out2 != 0, keep)
    torch.manual_seed(11)
    out3 = ops.rowop(x, None, False, None, gam, bet, 1e-5, False, p, torch.float32)
    assert torch.equal(out3, out.detach())
    # backward uses the same mask
    dy = torch.randn(rows, d, generator=g).to(dev())
    (out * dy).sum().backward()
    xb = x.clone().requires_grad_(True)
    (F.layer_norm(xb, (d,), gam, bet, 1e-5) * keep.float() * scale * dy).sum().backward()
    assert (xa.grad - xb.grad).abs().max().item() <= 1e-4 * xb.grad.abs().max().item() + 1e-6


def test_rowop_unsupported_width_composes_aten_ops():
    from allset_b200 import ops
    x = torch.randn(100, 70, device=dev(), requires_grad=True)
    b = torch.randn(70, device=dev())
    out = ops.rowop(x, b, True)
    torch.testing.assert_close(out, F.relu(x + b))
    out.sum().backward()
    assert x.grad is not None


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_linear_nb_forward_backward(dtype):
    from allset_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, 128, generator=g).to(dev())
    w = (torch.randn(64, 128, generator=g) / 11).to(dev())
    dy = torch.randn(5000, 64, generator=g).to(dev())
    xa = x.to(dtype).clone().requires_grad_(True)
    wa = w.clone().requires_grad_(True)
    y = ops.linear_nb(xa, wa)
    assert y.dtype == dtype
    xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    ref = xb @ wb.t()
    lo = dtype == torch.bfloat16
    assert (y.float() - ref).abs().max().item() <= (2e-2 if lo else 1e-4) * ref.abs().max().item()
    (y.float() * dy).sum().backward()
    (ref * dy).sum().backward()
    assert wa.grad.dtype == torch.float32
    assert_grad_close(wa.grad, wb.grad, 'w', rel=1e-2 if lo else 1e-4)
    assert (xa.grad.float() - xb.grad).abs().max().item() <= (2e-2 if lo else 1e-4) * xb.grad.abs().max().item()


def _bf16_mode_grads(rec):
    model, data = _build(rec, agg_dtype=torch.bfloat16)
    out = model(data)
    assert out.dtype == torch.float32
    (out * rec['grad_logits'].to(dev())).sum().backward()
    grads = dict((k, p.grad) for k, p in model.named_parameters() if p.grad is not None)
    assert all(g.dtype == torch.float32 for g in grads.values())
    return out.detach().cpu(), dict((k, g.float().cpu().reshape(-1)) for k, g in grads.items())


@pytest.mark.parametrize('name', ['cora_alldeepsets.pt', 'citeseer_allsettransformer.pt'])
def test_bf16_mode_training_gradients_vs_reference(name, monkeypatch):
    """Autograd through the bf16 training chain (bf16 GEMM operands on the hand-written tcgen05 Linear kernels, bf16
    activations and gathered rows, rowop forward / backward) on the real models.

    Two bars.  (1) Against the SAME recipe with the three GEMMs of every Linear handed to cuBLAS (`ops.TC_LINEAR = False`):
    the kernels compute the same products with the same fp32 accumulation, so logits and every parameter gradient must
    agree to 2 % of the gradient's norm -- this is the parity statement for the kernels.  (2) Against the reference's fp32
    logits and gradients: what the MODE costs.  bf16 operands perturb every activation by 2^-9 and the LayerNorm chain
    amplifies a perturbation ~3x per half layer (profiles/r02_bf16_error_budget.md: storing ONLY the gathered rows in bf16
    already costs 1.1e-2 of the logit scale on cora); measured with an fp64 evaluation of the oracle as the arbiter
    (scripts/diag_split.py, gpurun r2j) the cuBLAS-backed recipe itself is 3.6e-2 off in the logits and up to 30 % off in
    the L2 norm of individual LayerNorm / bias gradients on cora (d = 64, 1433 -> 64 first layer).  So the bar against the
    reference is the one a mixed-precision recipe can meet: logits within 6e-2 of the scale, every non-vanishing
    gradient within 40 % of its own L2 norm with cosine >= 0.93 (weights: >= 0.97)."""
    from allset_b200 import ops
    monkeypatch.setattr(ops, 'FUSED_DENSE_MIN_ROWS', 0)
    rec = load_golden(name)
    out, grads = _bf16_mode_grads(rec)
    monkeypatch.setattr(ops, 'TC_LINEAR', False)
    out_lib, grads_lib = _bf16_mode_grads(rec)
    monkeypatch.setattr(ops, 'TC_LINEAR', True)
    scale = rec['logits'].abs().max().item()
    big = max(g.norm().item() for g in rec['grads'].values())
    # (1) tcgen05 kernels vs cuBLAS inside the same bf16 recipe
    assert (out - out_lib).abs().max().item() <= 5e-3 * max(scale, 1.0)
    for k, g in grads_lib.items():
        err = (grads[k] - g).norm().item()
        assert err <= 2e-2 * g.norm().item() + 2e-3 * big, '%s: tcgen05 vs cuBLAS |err| %.3e vs |g| %.3e' % (k, err, g.norm().item())
    # (2) the mode vs the reference's fp32 arithmetic
    assert (out - rec['logits']).abs().max().item() <= 6e-2 * max(scale, 1.0)
    for k, g in rec['grads'].items():
        mine, ref = grads[k], g.reshape(-1)
        err = (mine - ref).norm().item()
        # a few parameters have (near-)vanishing gradients by symmetry (a bias in front of a softmax shifts every score of
        # a head alike): their error is measured against the model's gradient scale, not their own
        assert err <= 0.4 * ref.norm().item() + 0.01 * big, '%s: |err| %.3e vs |ref| %.3e (largest %.3e)' % (k, err, ref.norm().item(), big)
        if ref.norm().item() > 0.05 * big:
            cos = torch.dot(mine, ref).item() / (mine.norm().item() * ref.norm().item() + 1e-30)
            assert cos >= (0.97 if ref_is_matrix(g) else 0.93), '%s: cosine %.4f' % (k, cos)


def ref_is_matrix(g):
    return g.dim() >= 2 and min(g.shape[-2:]) > 1


def _train_losses(pma, agg, dropout, steps, min_rows):
    import allset_oracle as O
    from allset_b200 import ops, synthetic, preprocessing as P
    from types import SimpleNamespace
    n, m, d = 20000, 4000, 128
    v2e = synthetic.poisson_hypergraph(n, m, 8, seed=3, device=dev())
    ei, tot = P.add_self_loops(v2e, n, m)
    norm = P.norm_construction(ei)
    x = synthetic.features(n, d, torch.float32, device=dev())
    y = (x[:, :10].argmax(dim=1)).long()                                   # learnable from the features
    args = O.config_namespace(num_features=d, num_classes=10, MLP_hidden=d, Classifier_hidden=d, heads=8 if pma else 1,
                              All_num_layers=1, Classifier_num_layers=1, PMA=pma, aggregate='add', dropout=dropout)
    old = ops.FUSED_DENSE_MIN_ROWS
    ops.FUSED_DENSE_MIN_ROWS = min_rows
    try:
        torch.manual_seed(0)
        model = ab().SetGNN(args, agg_dtype=agg).to(dev()).train()
        opt = torch.optim.Adam(model.parameters(), lr=3e-3)
        data = SimpleNamespace(x=x, edge_index=ei.clone(), norm=norm)
        losses = []
        for _ in range(steps):
            opt.zero_grad(set_to_none=True)
            loss = F.nll_loss(torch.log_softmax(model(data), dim=1), y)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        return losses
    finally:
        ops.FUSED_DENSE_MIN_ROWS = old


@pytest.mark.parametrize('pma', [False, True])
def test_training_trajectory_matches_the_aten_path_without_dropout(pma):
    """dropout = 0 makes training deterministic: 12 Adam steps through the fused chain (rowop forward / backward kernels,
    bias-free Linears) follow the loss trajectory of the plain ATen modules (nn.Linear / F.relu / nn.LayerNorm with
    torch autograd) -- fp32 chain within 2e-3, bf16 mode within 5e-2 of the fp32 losses."""
    import torch.nn as nn
    # dropout 0 everywhere except SetGNN's hard-wired input dropout (p = 0.2, reference src/models.py:473): seed it alike
    ref = _train_losses(pma, None, 0.0, 12, min_rows=1 << 40)              # ATen dense path
    chain = _train_losses(pma, None, 0.0, 12, min_rows=0)
    bf16 = _train_losses(pma, torch.bfloat16, 0.0, 12, min_rows=0)
    assert all(l == l for l in ref + chain + bf16)
    for a, b in zip(ref, chain):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1.0), (ref, chain)
    for a, b in zip(ref, bf16):                    # (bf16 mode draws SetGNN's input dropout from the counter hash: another mask)
        assert abs(a - b) <= 8e-2 * max(abs(a), 1.0), (ref, bf16)
    assert ref[-1] < ref[0] and chain[-1] < chain[0] and bf16[-1] < bf16[0]


@pytest.mark.parametrize('pma', [False, True])
@pytest.mark.parametrize('agg', [None, torch.bfloat16])
def test_training_step_with_dropout_runs(pma, agg):
    """train() mode with dropout 0.5 through the fused chain (counter-based dropout regenerated by the backward kernel):
    finite losses that go down."""
    losses = _train_losses(pma, agg, 0.5, 40, min_rows=0)
    assert all(l == l and l < 1e4 for l in losses), losses
    assert min(losses[-5:]) < losses[0], (agg, losses[0], losses[-5:])
