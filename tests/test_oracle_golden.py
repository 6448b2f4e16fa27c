"""The oracle (oracle/allset_oracle.py) against golden vectors recorded from the reference's own modules
(oracle/make_golden.py).  This is what pins the oracle: every restated function is checked against outputs of
/root/reference/src/layers.py + models.py run under the third-party shims."""
import pytest
import torch
import torch.nn.functional as F

import allset_oracle as O
from conftest import golden_x, load_golden

TOL = dict(rtol=1e-5, atol=1e-5)


def _setgnn(rec, x):
    a = rec['args']
    return O.setgnn(rec['state_dict'], x, rec['edge_index'], rec['norm'], PMA=a['PMA'], heads=a['heads'],
                    aggregate=a['aggregate'], dropout=a['dropout'], GPR=a['GPR'], LearnMask=a['LearnMask'], training=False)


def _check_setgnn(rec):
    x = golden_x(rec).clone().requires_grad_(True)
    logits, taps = _setgnn(rec, x)
    torch.testing.assert_close(logits, rec['logits'], **TOL)
    assert len(taps) == len(rec['taps'])
    s = rec['tap_stride']
    for mine, ref in zip(taps, rec['taps']):
        torch.testing.assert_close(mine[::s], F.relu(ref), **TOL)
    (logits * rec['grad_logits']).sum().backward()
    torch.testing.assert_close(x.grad.sum(dim=1), rec['grad_x_rowsum'], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('name', ['cora_alldeepsets.pt', 'citeseer_allsettransformer.pt'])
def test_real_datasets(name):
    rec = load_golden(name)
    _check_setgnn(rec)
    # the reference zero-bases hyperedge ids in place (models.py:453-454)
    assert int(rec['edge_index_after'][1].min()) == 0
    assert torch.equal(rec['edge_index_after'][1], rec['edge_index'][1] - rec['edge_index'][1].min())


def test_real_dataset_shapes():
    cora = load_golden('cora_alldeepsets.pt')
    assert cora['n_nodes'] == 2708 and cora['edge_index'].shape == (2, 7494)
    assert O.implied_rows(cora['edge_index'][1] - cora['edge_index'][1].min()) == 4287
    assert cora['norm'].dtype == torch.int64 and bool((cora['norm'] == 1).all())
    cs = load_golden('citeseer_allsettransformer.pt')
    assert cs['n_nodes'] == 3312 and cs['edge_index'].shape == (2, 6765)
    assert O.implied_rows(cs['edge_index'][1] - cs['edge_index'][1].min()) == 4391


@pytest.mark.parametrize('idx', range(12))
def test_setgnn_variants(idx):
    _check_setgnn(load_golden('setgnn_variants.pt')[idx])


@pytest.mark.parametrize('idx', range(12))
def test_layers(idx):
    rec = load_golden('layers_small.pt')[idx]
    x = rec['x'].clone().requires_grad_(True)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and 'running_' not in k)
              for k, v in rec['state_dict'].items()}
    src, tgt = rec['edge_index'][0], rec['edge_index'][1]
    if rec['kind'] == 'pma':
        out, alpha = O.pma(params, '', x, src, tgt, rec['extra']['heads'])
        torch.testing.assert_close(alpha, rec['alpha'], **TOL)
    else:
        out = O.half_nlh_conv(params, '', x, src, tgt, rec['norm'], rec['aggr'], attention=False)
    assert out.shape[0] == rec['extra']['n_tgt']
    torch.testing.assert_close(out, rec['out'], **TOL)
    (out * rec['grad_out']).sum().backward()
    torch.testing.assert_close(x.grad, rec['grad_x'], rtol=1e-4, atol=1e-5)
    for k, g in rec['grads'].items():
        torch.testing.assert_close(params[k].grad, g, rtol=1e-4, atol=1e-5)


# ---- properties of the restated third-party primitives --------------------------------------
def test_scatter_matches_dense_incidence():
    g = torch.Generator().manual_seed(0)
    n_src, n_tgt, nnz = 17, 9, 60
    src = torch.randint(0, n_src, (nnz,), generator=g)
    tgt = torch.randint(0, n_tgt - 1, (nnz,), generator=g)
    tgt[0] = n_tgt - 1
    x = torch.randn(n_src, 5, generator=g, dtype=torch.float64)
    w = torch.rand(nnz, generator=g, dtype=torch.float64)
    Hm = torch.zeros(n_tgt, n_src, dtype=torch.float64)
    Hm.index_put_((tgt, src), w, accumulate=True)
    torch.testing.assert_close(O.aggregate_sum_mean(x, src, tgt, w, 'sum'), Hm @ x)
    cnt = torch.bincount(tgt, minlength=n_tgt).clamp(min=1).double()
    torch.testing.assert_close(O.aggregate_sum_mean(x, src, tgt, w, 'mean'), (Hm @ x) / cnt[:, None])


def test_output_rows_follow_index_max():
    x = torch.randn(10, 3)
    src = torch.tensor([0, 1, 2]); tgt = torch.tensor([0, 4, 2])
    assert O.aggregate_sum_mean(x, src, tgt, None, 'sum').shape == (5, 3)          # rows 1, 3 are interior zeros
    assert bool((O.aggregate_sum_mean(x, src, tgt, None, 'mean')[[1, 3]] == 0).all())
    assert O.aggregate_sum_mean(x, src[:0], tgt[:0], None, 'sum').shape == (0, 3)


def test_softmax_rows_sum_to_one_and_empty_segments():
    g = torch.Generator().manual_seed(1)
    score = torch.randn(40, 4, generator=g) * 5
    idx = torch.randint(0, 6, (40,), generator=g); idx[idx == 3] = 2; idx[0] = 6
    a = O.segment_softmax(score, idx)
    s = O.scatter_rows(a, idx, 'sum')
    present = torch.bincount(idx, minlength=7) > 0
    torch.testing.assert_close(s[present], torch.ones_like(s[present]), rtol=1e-6, atol=1e-6)
    assert bool((s[~present] == 0).all())


def test_pma_empty_segment_is_seed():
    g = torch.Generator().manual_seed(2)
    v = torch.randn(6, 2, 3, generator=g); sc = torch.randn(6, 2, generator=g); seed = torch.randn(1, 2, 3, generator=g)
    src = torch.tensor([0, 1, 2, 3]); tgt = torch.tensor([0, 0, 2, 2])
    out, _ = O.aggregate_pma(v, sc, seed, src, tgt)
    torch.testing.assert_close(out[1], seed[0])
