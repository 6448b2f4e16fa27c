"""allset_b200.preprocessing against fixtures recorded from the reference's own ExtractV2E / Add_Self_Loops /
norm_contruction (oracle/make_golden.py::preprocessing_cases).

Index work, compared exactly -- up to the one thing the reference leaves unspecified: it orders incidences with
`torch.sort(edge_index[0])` (preprocessing.py:398,446), which is NOT stable, so the order of the hyperedges of one node
is whatever the sort implementation of the running torch build produces (it differs between CPU and CUDA and between
versions).  Row 0 (the node ids) must match element for element; (node, hyperedge) pairs must match as a multiset, i.e.
after a canonical lexicographic order; per-incidence norms are compared in that same canonical order."""
import pytest
import torch

from allset_b200 import preprocessing as P
from conftest import load_golden

CASES = range(5)


def canon(ei, *per_incidence):
    """Lexicographic (node, hyperedge) order; returns the permuted list and the permuted per-incidence tensors."""
    key = ei[0] * (int(ei[1].max()) + 1) + ei[1]
    order = torch.sort(key, stable=True)[1]
    return (ei[:, order],) + tuple(t[order] for t in per_incidence)


def same_incidences(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a[0], b[0]) and torch.equal(canon(a)[0], canon(b)[0])


@pytest.mark.parametrize('i', CASES)
def test_matches_reference_cpu(i):
    c = load_golden('preprocessing.pt')[i]
    v2e = P.extract_v2e(c['raw'], c['n_x'], c['num_hyperedges'])
    assert v2e.dtype == torch.int64 and same_incidences(v2e, c['v2e'])
    ei, tot = P.add_self_loops(v2e, c['n_x'], c['num_hyperedges'])
    assert tot == c['totedges'] and same_incidences(ei, c['with_loops'])
    ones = P.norm_construction(ei, 'all_one')
    assert ones.dtype == c['norm_all_one'].dtype and torch.equal(ones, c['norm_all_one'])
    sym = P.norm_construction(ei, 'deg_half_sym')
    assert sym.dtype == c['norm_deg_half_sym'].dtype
    torch.testing.assert_close(canon(ei, sym)[1], canon(c['with_loops'], c['norm_deg_half_sym'])[1], rtol=1e-6, atol=0)
    torch.testing.assert_close(canon(v2e, P.norm_construction(v2e, 'deg_half_sym'))[1],
                               canon(c['v2e'], c['norm_deg_half_sym_noloop'])[1], rtol=1e-6, atol=0)
    ei2, norm2, tot2 = P.preprocess(c['raw'], c['n_x'], c['num_hyperedges'])
    assert same_incidences(ei2, c['with_loops']) and tot2 == c['totedges'] and torch.equal(norm2, c['norm_all_one'])


def test_real_cora_numbers():
    c = load_golden('preprocessing.pt')[0]
    assert c['name'] == 'cora' and c['n_x'] == 2708 and c['num_hyperedges'] == 1579
    ei, tot = P.add_self_loops(P.extract_v2e(c['raw'], 2708, 1579), 2708, 1579)
    assert ei.shape == (2, 7494) and tot == 4287
    assert bool((ei[0, 1:] >= ei[0, :-1]).all()) and int(ei[1].min()) == 2708


def test_error_paths():
    c = load_golden('preprocessing.pt')[2]
    with pytest.raises(ValueError, match='does not match'):
        P.extract_v2e(c['raw'], c['n_x'], c['num_hyperedges'] + 1)
    with pytest.raises(ValueError, match='does not match'):
        P.add_self_loops(c['v2e'], c['n_x'], c['num_hyperedges'] + 3)
    with pytest.raises(ValueError):
        P.norm_construction(c['v2e'], 'bogus')


@pytest.mark.gpu
@pytest.mark.parametrize('i', CASES)
def test_matches_reference_on_gpu(i):
    c = load_golden('preprocessing.pt')[i]
    dev = torch.device('cuda:0')
    ei, norm, tot = P.preprocess(c['raw'].to(dev), c['n_x'], c['num_hyperedges'], normtype='deg_half_sym')
    assert ei.is_cuda and same_incidences(ei.cpu(), c['with_loops']) and tot == c['totedges']
    torch.testing.assert_close(canon(ei.cpu(), norm.cpu())[1], canon(c['with_loops'], c['norm_deg_half_sym'])[1], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_large_graph_on_gpu_feeds_setgnn_layout():
    """1M-vertex star expansion: seconds of Python loops in the reference, milliseconds here; invariants only."""
    from allset_b200 import synthetic
    dev = torch.device('cuda:0')
    n, m = 1_000_000, 200_000
    v2e = synthetic.poisson_hypergraph(n, m, 20, seed=3, device=dev)
    raw = torch.cat([v2e, v2e.flip(0)], dim=1)
    ei, norm, tot = P.preprocess(raw, n, m)
    assert bool((ei[0, 1:] >= ei[0, :-1]).all())
    sizes = torch.bincount(ei[1] - n)
    assert tot == sizes.numel() and int(ei[1].max()) == n + tot - 1
    # every node is now the sole member of exactly one size-1 hyperedge or already was
    single = sizes[ei[1] - n] == 1
    covered = torch.zeros(n, dtype=torch.bool, device=dev)
    covered[ei[0][single]] = True
    assert bool(covered.all())
    assert norm.dtype == torch.int64 and bool((norm == 1).all())


# ---------------------------------------------------------------------------------------------------------
# expand_edge_index (--exclude_self): fixtures from the reference's own function (oracle/make_golden_expand.py)
# ---------------------------------------------------------------------------------------------------------
def _expand_cases():
    return load_golden('expand_edge_index.pt')


@pytest.mark.parametrize('i', range(5))
def test_expand_edge_index_matches_reference(i):
    c = _expand_cases()[i]
    mine = P.expand_edge_index(c['edge_index'], c['n_x'], c['n_he'], edge_th=c['edge_th'])
    ref = c['expanded']
    # same node row element for element, same (node, new hyperedge id) pairs; the order of one node's hyperedges is left
    # to an unstable argsort by the reference (preprocessing.py:141)
    assert mine.dtype == ref.dtype and same_incidences(mine, ref), c['name']


def test_expand_edge_index_counts_and_semantics():
    c = _expand_cases()[1]
    ei, n = c['edge_index'], c['n_x']
    out = P.expand_edge_index(ei, n, c['n_he'])
    sizes = torch.bincount(ei[1] - n)
    sizes = sizes[sizes > 0]
    assert out.shape[1] == int((sizes * (sizes - 1)).clamp(min=0).sum() + (sizes == 1).sum())   # s(s-1), singletons kept
    assert int(out[1].min()) == n and int(out[1].max()) == n + int(sizes.sum()) - 1             # one new id per member
    new_sizes = torch.bincount(out[1] - n)
    assert bool(((new_sizes >= 1)).all())
    # through preprocess(): ExtractV2E -> Add_Self_Loops -> expand
    g = load_golden('preprocessing.pt')[2]
    ei2, norm2, tot2 = P.preprocess(g['raw'], g['n_x'], g['num_hyperedges'], exclude_self=True)
    assert ei2.shape[1] == norm2.numel() and tot2 == int(ei2[1].max()) - g['n_x'] + 1
    assert bool((ei2[0, 1:] >= ei2[0, :-1]).all())


@pytest.mark.gpu
def test_expand_edge_index_on_gpu_large():
    from allset_b200 import synthetic
    dev = torch.device('cuda:0')
    c = _expand_cases()[0]
    mine = P.expand_edge_index(c['edge_index'].to(dev), c['n_x'], c['n_he'])
    assert mine.is_cuda and same_incidences(mine.cpu(), c['expanded'])
    n, m = 300_000, 60_000
    v2e = synthetic.poisson_hypergraph(n, m, 10, seed=5, device=dev)
    out = P.expand_edge_index(v2e, n, m)
    sizes = torch.bincount(v2e[1] - n, minlength=m)
    assert out.shape[1] == int((sizes * (sizes - 1)).sum() + (sizes == 1).sum())
    assert bool((out[0, 1:] >= out[0, :-1]).all())
