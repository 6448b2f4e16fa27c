"""allset_b200.uni (UniGCNII on the segmented-reduce kernels) against fixtures recorded from the
reference's own UniGCNII / UniGCNIIConv (oracle/make_golden_uni.py; reference src/models.py:909-995)."""
from types import SimpleNamespace

import pytest
import torch

from conftest import load_golden


def _build(c, device):
    import allset_b200
    args = SimpleNamespace(UniGNN_degV=c['degV'].to(device), UniGNN_degE=c['degE'].to(device), UniGNN_use_norm=c['use_norm'])
    m = allset_b200.UniGCNII(args, V=c['V'].to(device), E=c['E'].to(device), **c['ctor'])
    missing = m.load_state_dict(c['state_dict'], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(device).eval()


@pytest.mark.parametrize('i', range(3))
def test_state_dict_and_api_match_reference_cpu(i):
    c = load_golden('unigcnii.pt')[i]
    m = _build(c, 'cpu')
    assert list(m.state_dict().keys()) == list(c['state_dict'].keys())
    assert len(m.reg_params) == c['ctor']['nlayer'] and len(m.non_reg_params) == 4
    with pytest.raises(RuntimeError, match='CUDA only'):
        m(SimpleNamespace(x=c['x']))                    # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize('i', range(3))
def test_unigcnii_matches_reference_gpu(i):
    c = load_golden('unigcnii.pt')[i]
    dev = torch.device('cuda:0')
    m = _build(c, dev)
    x = c['x'].to(dev).requires_grad_(True)
    out = m(SimpleNamespace(x=x))
    torch.testing.assert_close(out.detach().cpu(), c['logits'], rtol=1e-4, atol=1e-4)
    (out * c['grad_logits'].to(dev)).sum().backward()
    torch.testing.assert_close(x.grad.cpu(), c['grad_x'], rtol=1e-3, atol=1e-4)
    for k, p in m.named_parameters():
        ref = c['grads'][k]
        err = (p.grad.cpu() - ref).abs().max().item()
        assert err <= 1e-3 * ref.abs().max().item() + 1e-5, (k, err)
    conv = m.convs[1]
    h = c['conv_in'].to(dev)
    with torch.no_grad():
        got = conv(h, m.V, m.E, 0.1, 0.4, 0.5 * h)
    torch.testing.assert_close(got.cpu(), c['conv_out'], rtol=1e-4, atol=1e-4)
    # second forward reuses the cached incidence
    with torch.no_grad():
        assert torch.equal(m(SimpleNamespace(x=x.detach())), out.detach())


@pytest.mark.gpu
def test_unigcnii_large_graph_runs_on_stream_kernels():
    """config-3-sized incidence list: both reductions take the stream kernels; result vs a dense-free torch restatement."""
    import allset_b200
    from allset_b200 import synthetic
    dev = torch.device('cuda:0')
    n, m_e, d = 300_000, 80_000, 128
    ei = synthetic.poisson_hypergraph(n, m_e, 12, seed=9, device=dev)
    V, E = ei[0].contiguous(), (ei[1] - n).contiguous()
    degV = torch.bincount(V, minlength=n).view(-1, 1).float()
    cnt = torch.bincount(E, minlength=m_e).view(-1, 1).float()
    degE = (torch.zeros(m_e, 1, device=dev).index_add_(0, E, degV[V]) / cnt.clamp(min=1)).pow(-0.5)
    degV = degV.pow(-0.5)
    degV[torch.isinf(degV)] = 1
    args = SimpleNamespace(UniGNN_degV=degV, UniGNN_degE=degE, UniGNN_use_norm=False)
    conv = allset_b200.UniGCNIIConv(args, d, d).to(dev)
    x = torch.randn(n, d, device=dev)
    with torch.no_grad():
        got = conv(x, V, E, 0.1, 0.3, x)
        xe = torch.zeros(m_e, d, device=dev).index_add_(0, E, x[V]) / cnt.clamp(min=1) * degE
        xv = torch.zeros(n, d, device=dev).index_add_(0, V, xe[E]) * degV
        xi = 0.9 * xv + 0.1 * x
        ref = 0.7 * xi + 0.3 * conv.W(xi)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
