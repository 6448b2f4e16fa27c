"""The hand-written tcgen05 Linear kernels (allset_linear_fwd, allset_linear_wgrad) that carry every square nn.Linear of
the path (reference src/layers.py:575 MLP.lins, :128-130 PMA.lin_K / lin_V, :76-80 rFF) in BOTH modes:

  * fp32 rows -> split precision (forward / input gradient: three bf16 terms per operand, six tensor-core products,
    fp32 accumulate -- measured against an fp64 product at the level of an fp32 SGEMM, 5e-6 of the output scale; weight
    gradient: two terms, 3e-5), far inside the reference's fp32 bar of 1e-4;
  * bf16 rows -> bf16 operands: the 1e-2 bar.

Forward with the LayerNorm prologue / bias / ReLU, the input gradient (weight read transposed), the weight gradient
(MN-major operands, partial per CTA + ordered reduce), ragged row counts, and autograd through ops.linear_nb."""
import pytest
import torch
import torch.nn.functional as F

from test_gpu_parity import dev

pytestmark = pytest.mark.gpu

ROWS = [1, 63, 64, 129, 5000, 40007]


def _mk(rows, d, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, d, generator=g) * 1.5 + 0.25
    w = (torch.rand(d, d, generator=g) * 2 - 1) / d ** 0.5
    b = torch.randn(d, generator=g) * 0.3
    return x.to(dtype).to(dev()), w.to(dev()), b.to(dev())


def _status():
    return torch.zeros(1, dtype=torch.int32, device=dev())


@pytest.mark.parametrize('d', [64, 128])
@pytest.mark.parametrize('rows', ROWS + [300001])
def test_linear_fwd_split_precision_vs_fp64(d, rows):
    from allset_b200 import _lib
    x, w, b = _mk(rows, d, 11 + d + rows)
    st = _status()
    out = _lib.linear_fwd(x, w, b, relu=False, status=st)
    assert out.dtype == torch.float32 and int(st.item()) == 0
    ref = (x.double() @ w.double().t() + b.double())
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item()
    assert err <= 5e-6 * scale, 'split-precision forward: %.3e of the scale' % (err / scale)
    # and it is at least as good as the fp32 SGEMM it replaces is far from fp64 by construction; the 1e-4 bar of the path:
    assert (out - (x @ w.t() + b)).abs().max().item() <= 1e-4 * scale


@pytest.mark.parametrize('d', [64, 128])
def test_linear_fwd_split_ln_bias_relu(d):
    from allset_b200 import _lib
    rows = 7001
    x, w, b = _mk(rows, d, 5 + d)
    g = torch.Generator().manual_seed(d)
    gam = (torch.rand(d, generator=g) + 0.5).to(dev())
    bet = (torch.randn(d, generator=g) * 0.2).to(dev())
    st = _status()
    out = _lib.linear_fwd(x, w, b, ln=(gam, bet, 1e-5), relu=True, status=st)
    assert int(st.item()) == 0
    ref = F.relu(F.linear(F.layer_norm(x.double(), (d,), gam.double(), bet.double(), 1e-5), w.double(), b.double()))
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 1e-5 * scale
    out2 = _lib.linear_fwd(x, w, None, ln=(gam, None, 1e-5), relu=False)
    ref2 = F.linear(F.layer_norm(x.double(), (d,), gam.double(), None, 1e-5), w.double())
    assert (out2.double() - ref2).abs().max().item() <= 1e-5 * ref2.abs().max().item()


@pytest.mark.parametrize('d', [64, 128])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_linear_fwd_transposed_is_the_input_gradient(d, dtype):
    from allset_b200 import _lib
    rows = 9000 + d
    dy, w, _ = _mk(rows, d, 3 + d, dtype)
    st = _status()
    dx = _lib.linear_fwd(dy, w, transposed=True, status=st)
    assert dx.dtype == dtype and int(st.item()) == 0
    ref = dy.double() @ w.double()
    scale = ref.abs().max().item()
    tol = 5e-6 if dtype == torch.float32 else 1e-2
    assert (dx.double() - ref).abs().max().item() <= tol * scale


@pytest.mark.parametrize('d', [64, 128])
@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
def test_linear_fwd_bf16_rows(d, out_dtype):
    from allset_b200 import _lib
    rows = 12345
    x, w, b = _mk(rows, d, 9 + d, torch.bfloat16)
    out = _lib.linear_fwd(x, w, b, relu=True, out_dtype=out_dtype)
    assert out.dtype == out_dtype
    ref = F.relu(x.float() @ w.t() + b)
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize('d', [64, 128])
@pytest.mark.parametrize('rows', ROWS + [300001])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_linear_wgrad_vs_fp64(d, rows, dtype):
    from allset_b200 import _lib
    g = torch.Generator().manual_seed(rows + d)
    dy = (torch.randn(rows, d, generator=g)).to(dtype).to(dev())
    x = (torch.randn(rows, d, generator=g) * 1.5 + 0.25).to(dtype).to(dev())
    st = _status()
    dw = _lib.linear_wgrad(dy, x, status=st)
    assert dw.dtype == torch.float32 and tuple(dw.shape) == (d, d) and int(st.item()) == 0
    ref = dy.double().t() @ x.double()
    # a weight gradient is a sum over `rows` signed products: its error is measured against sqrt(rows) * |dy| * |x| (the
    # magnitude a random sum reaches), which is also the scale of ref for these inputs
    scale = max(ref.abs().max().item(), 1e-30)
    err = (dw.double() - ref).abs().max().item()
    if dtype == torch.float32:
        assert err <= 3e-5 * scale, 'split-precision wgrad: %.3e of the scale' % (err / scale)
    else:
        assert err <= 1e-3 * scale, 'bf16 wgrad (exact products of bf16 inputs, fp32 accumulate): %.3e' % (err / scale)
    # deterministic: ordered partial sums, no atomics
    assert torch.equal(dw, _lib.linear_wgrad(dy, x))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('d', [64, 128])
def test_linear_nb_square_autograd_on_tcgen05(dtype, d, monkeypatch):
    from allset_b200 import ops
    monkeypatch.setattr(ops, 'FUSED_DENSE_MIN_ROWS', 0)
    rows = 20011
    x, w, _ = _mk(rows, d, 21 + d)
    g = torch.Generator().manual_seed(2)
    dy = torch.randn(rows, d, generator=g).to(dev())
    xa = x.to(dtype).clone().requires_grad_(True)
    wa = w.clone().requires_grad_(True)
    assert ops.tc_linear_ok(xa, wa)
    y = ops.linear_nb(xa, wa)
    assert y.dtype == dtype
    xb, wb = x.to(dtype).double().requires_grad_(True), w.double().requires_grad_(True)
    ref = xb @ wb.t()
    lo = dtype == torch.bfloat16
    assert (y.double() - ref).abs().max().item() <= (1e-2 if lo else 5e-6) * ref.abs().max().item()
    (y.float() * dy.to(dtype).float()).sum().backward()
    (ref * dy.to(dtype).double()).sum().backward()
    assert wa.grad.dtype == torch.float32 and xa.grad.dtype == dtype
    assert (wa.grad.double() - wb.grad).abs().max().item() <= (2e-3 if lo else 3e-5) * wb.grad.abs().max().item()
    assert (xa.grad.double() - xb.grad).abs().max().item() <= (1e-2 if lo else 5e-6) * xb.grad.abs().max().item()


def test_fp32_mode_mlp_eval_chain_matches_reference_arithmetic(monkeypatch):
    """MLP.forward in fp32 mode without autograd = one split-precision launch per Linear (LayerNorm prologue, bias, ReLU
    fused) -- against the module's own ATen path (reference src/layers.py:571-579) at the fp32 bar."""
    from allset_b200 import ops
    from allset_b200.layers import MLP
    torch.manual_seed(0)
    d = 128
    mlp = MLP(d, d, d, 2, dropout=0.5, Normalization='ln', InputNorm=True).to(dev()).eval()
    with torch.no_grad():
        for n in mlp.normalizations:
            n.weight.uniform_(0.5, 1.5)
            n.bias.normal_(0, 0.2)
    x = torch.randn(30000, d, device=dev()) * 2 + 0.3
    with torch.no_grad():
        monkeypatch.setattr(ops, 'TC_LINEAR', False)
        monkeypatch.setattr(ops, 'FUSED_DENSE_MIN_ROWS', 1 << 60)
        ref = F.relu(mlp(x))
        monkeypatch.setattr(ops, 'TC_LINEAR', True)
        monkeypatch.setattr(ops, 'FUSED_DENSE_MIN_ROWS', 0)
        out = mlp(x, final_relu=True)
    assert out.dtype == torch.float32
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


@pytest.mark.parametrize('pma', [False, True])
def test_fp32_mode_model_eval_on_tcgen05_matches_the_aten_path(pma, monkeypatch):
    """Whole SetGNN forward in fp32 mode at a size where every square Linear runs in split precision on tcgen05 (MLP chain
    with fused LayerNorm / bias / ReLU, PMA.lin_V and rFF through ops.linear_bias_act) against the same model with the
    kernels switched off (cuBLAS SGEMM + glue): the fp32 bar, 1e-4 of the logit scale."""
    import allset_oracle as O
    from types import SimpleNamespace
    from allset_b200 import ops, synthetic, preprocessing as P
    import allset_b200
    n, m, d = 30000, 6000, 128
    v2e = synthetic.poisson_hypergraph(n, m, 8, seed=5, device=dev())
    ei, tot = P.add_self_loops(v2e, n, m)
    norm = P.norm_construction(ei)
    x = synthetic.features(n, d, torch.float32, device=dev())
    args = O.config_namespace(num_features=d, num_classes=7, MLP_hidden=d, Classifier_hidden=d, heads=8 if pma else 1,
                              All_num_layers=2, Classifier_num_layers=1, PMA=pma, aggregate='add')
    torch.manual_seed(0)
    model = allset_b200.SetGNN(args).to(dev()).eval()
    with torch.no_grad():
        monkeypatch.setattr(ops, 'TC_LINEAR', False)
        ref = model(SimpleNamespace(x=x, edge_index=ei.clone(), norm=norm))
        monkeypatch.setattr(ops, 'TC_LINEAR', True)
        out = model(SimpleNamespace(x=x, edge_index=ei.clone(), norm=norm))
    assert (out - ref).abs().max().item() <= 1e-4 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize('d', [64, 128])
def test_linear_fwd_split_vs_the_oracle_restatement(d):
    """The kernel against the CPU restatement of its own arithmetic (oracle/allset_oracle.py::linear_split: three bf16 terms
    per operand, six exact products, fp32 accumulation): same terms, same products, only the order of the fp32 additions
    differs -- 1e-6 of the scale (the two are 2e-7 and 4e-7 away from fp64 themselves)."""
    import allset_oracle as O
    from allset_b200 import _lib
    x, w, b = _mk(3000, d, 77 + d)
    out = _lib.linear_fwd(x, w)
    ref = O.linear_split(x.cpu(), w.cpu(), 3)
    assert (out.cpu() - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
