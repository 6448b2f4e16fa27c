"""allset_b200.baselines (HCHA / HypergraphConv / HNHN / UniGNN family on the segmented-reduce and PMA kernels) against
fixtures recorded from the reference's OWN classes (oracle/make_golden_baselines.py; reference src/layers.py:233-494,
src/models.py:207-292,601-907): same state_dict keys, outputs and gradients within 1e-4."""
from types import SimpleNamespace

import pytest
import torch

from conftest import load_golden

CASES = load_golden('baselines.pt')


def _build(rec, device):
    from allset_b200 import baselines as B
    kind = rec['kind']
    if kind == 'HCHA':
        model = B.HCHA(SimpleNamespace(**rec['args']))
    elif kind == 'HypergraphConvAttention':
        model = B.HypergraphConv(**rec['ctor'])
    elif kind == 'HNHN':
        model = B.HNHN(SimpleNamespace(**rec['args']))
    else:
        args = SimpleNamespace(degV=rec['degV'].to(device), degE=rec['degE'].to(device), **rec['args'])
        model = B.UniGNN(args, V=rec['V'].to(device), E=rec['E'].to(device), **rec['ctor'])
    return model


@pytest.mark.parametrize('idx', range(len(CASES)))
def test_state_dict_keys_match_the_reference(idx):
    rec = CASES[idx]
    model = _build(rec, 'cpu')
    assert list(model.state_dict().keys()) == list(rec['state_dict'].keys())
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(rec['state_dict'][k].shape), k


@pytest.mark.gpu
@pytest.mark.parametrize('idx', range(len(CASES)))
def test_forward_and_gradients_match_the_reference(idx):
    rec = CASES[idx]
    dev = torch.device('cuda:0')
    model = _build(rec, dev)
    model.load_state_dict(rec['state_dict'], strict=True)
    model.to(dev).eval()
    x = rec['x'].to(dev).requires_grad_(True)
    kind = rec['kind']
    if kind == 'HCHA':
        out = model(SimpleNamespace(x=x, edge_index=rec['edge_index'].to(dev)))
    elif kind == 'HypergraphConvAttention':
        out = model(x, rec['edge_index'].to(dev))
    elif kind == 'HNHN':
        extra = {k: v.to(dev) for k, v in rec['extra'].items()}
        out = model(SimpleNamespace(x=x, edge_index=rec['edge_index'].to(dev), **extra))
    else:
        out = model(x)
    torch.testing.assert_close(out.detach().cpu(), rec['out'], rtol=1e-4, atol=1e-4)
    (out * rec['grad_out'].to(dev)).sum().backward()
    torch.testing.assert_close(x.grad.cpu(), rec['grad_x'], rtol=1e-3, atol=1e-4)
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(rec['grads']) <= set(grads)
    for k, g in rec['grads'].items():
        err = (grads[k].cpu() - g).abs().max().item()
        assert err <= 1e-3 * g.abs().max().item() + 1e-5, (k, err)
    # a second forward reuses the cached incidence
    with torch.no_grad():
        if kind == 'UniGNN':
            again = model(x)
            assert torch.equal(again, out.detach())
