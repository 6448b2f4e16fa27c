"""allset_b200.ingest: star expansion equal to the reference loader's output (the `raw` lists recorded from
load_citation_dataset in tests/golden/preprocessing.pt; the hypergraph pickles exist only in the dev container) and the
flat binary cache round trip."""
import io
import os
import pickle
import zipfile

import pytest
import torch

from allset_b200 import ingest
from conftest import load_golden

RAW_ZIP = os.path.join(os.environ.get('ALLSET_REFERENCE_ROOT', '/root/reference'), 'data', 'raw_data', 'AllSet_all_raw_data.zip')


@pytest.mark.skipif(not os.path.isfile(RAW_ZIP), reason='reference raw data not present')
@pytest.mark.parametrize('i,name', [(0, 'cora'), (1, 'citeseer')])
def test_star_expansion_equals_reference_loader(i, name):
    c = load_golden('preprocessing.pt')[i]
    with zipfile.ZipFile(RAW_ZIP) as z:
        hypergraph = pickle.load(io.BytesIO(z.read('AllSet_all_raw_data/cocitation/%s/hypergraph.pickle' % name)))
    ei, m = ingest.star_expansion(hypergraph, c['n_x'])
    assert m == c['num_hyperedges'] and ei.dtype == torch.int64
    assert torch.equal(ei, c['raw'])                    # bit-identical to load_citation_dataset + coalesce


def test_star_expansion_small_and_errors():
    ei, m = ingest.star_expansion([[0, 2], [1], [2, 0, 2]], 3)          # a duplicate member is coalesced away
    assert m == 3
    want = torch.tensor([[0, 0, 1, 2, 2, 3, 3, 4, 5, 5], [3, 5, 4, 3, 5, 0, 2, 1, 0, 2]])
    assert torch.equal(ei, want)
    with pytest.raises(ValueError):
        ingest.star_expansion([[0, 7]], 3)


def test_cache_round_trip(tmp_path):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100, 17, generator=g)
    xb = torch.randn(40, 8, generator=g).bfloat16()
    ei = torch.randint(0, 100, (2, 333), generator=g)
    y = torch.randint(0, 5, (100,), generator=g)
    p = str(tmp_path / 'g.allset')
    ingest.save_cache(p, x, ei, y, n_x=100, num_hyperedges=23, xb=xb, norm=torch.ones(333, dtype=torch.int64))
    d = ingest.load_cache(p)
    assert d.n_x == 100 and d.num_hyperedges == 23
    assert torch.equal(d.x, x) and torch.equal(d.edge_index, ei) and torch.equal(d.y, y)
    assert d.xb.dtype == torch.bfloat16 and torch.equal(d.xb, xb) and torch.equal(d.norm, torch.ones(333, dtype=torch.int64))
    assert os.path.getsize(p) % 1 == 0 and not os.path.exists(p + '.tmp')
    with open(str(tmp_path / 'bad'), 'wb') as f:
        f.write(b'nonsense')
    with pytest.raises(ValueError):
        ingest.load_cache(str(tmp_path / 'bad'))


@pytest.mark.gpu
def test_cache_to_device_feeds_preprocessing(tmp_path):
    from allset_b200 import preprocessing as P
    c = load_golden('preprocessing.pt')[0]
    p = str(tmp_path / 'cora.allset')
    ingest.save_cache(p, torch.zeros(c['n_x'], 4), c['raw'], None, n_x=c['n_x'], num_hyperedges=c['num_hyperedges'])
    d = ingest.load_cache(p, device='cuda:0', pin=True)
    assert d.edge_index.is_cuda
    ei, norm, tot = P.preprocess(d.edge_index, d.n_x, d.num_hyperedges)
    assert tot == c['totedges'] and ei.shape == c['with_loops'].shape


# --- the two text formats, against the reference's own loaders (dev container only: needs the raw zip + shims) ---------
def _extract(tmp_path, members):
    with zipfile.ZipFile(RAW_ZIP) as z:
        for src, dst in members:
            os.makedirs(os.path.dirname(str(tmp_path / dst)), exist_ok=True)
            with open(str(tmp_path / dst), 'wb') as f:
                f.write(z.read('AllSet_all_raw_data/' + src))


def _reference_loaders():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
    import ref_harness
    return ref_harness.load().loaders, ref_harness._quiet


@pytest.mark.skipif(not os.path.isfile(RAW_ZIP), reason='reference raw data not present')
def test_le_loader_equals_reference(tmp_path):
    _extract(tmp_path, [('zoo/zoo.content', 'zoo/zoo.content'), ('zoo/zoo.edges', 'zoo/zoo.edges')])
    loaders, quiet = _reference_loaders()
    with quiet():
        ref = loaders.load_LE_dataset(path=str(tmp_path), dataset='zoo')
    mine = ingest.load_le_dataset(str(tmp_path), 'zoo')
    assert mine.n_x == int(ref.n_x) and mine.num_hyperedges == int(ref.num_hyperedges)
    assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.y, ref.y)
    torch.testing.assert_close(mine.x, ref.x, rtol=0, atol=0)


@pytest.mark.skipif(not os.path.isfile(RAW_ZIP), reason='reference raw data not present')
def test_cornell_loader_equals_reference(tmp_path):
    name = 'house-committees'
    _extract(tmp_path, [('%s/node-labels-%s.txt' % (name, name), '%s/node-labels-%s.txt' % (name, name)),
                        ('%s/hyperedges-%s.txt' % (name, name), '%s/hyperedges-%s.txt' % (name, name))])
    loaders, quiet = _reference_loaders()
    with quiet():
        ref = loaders.load_cornell_dataset(path=str(tmp_path), dataset=name, feature_noise=0.0)
    mine = ingest.load_cornell_dataset(str(tmp_path), name, feature_noise=0.0)
    assert mine.n_x == int(ref.n_x) and mine.num_hyperedges == int(ref.num_hyperedges)
    assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.y, ref.y)
    torch.testing.assert_close(mine.x, ref.x, rtol=0, atol=0)            # noise 0: one-hot of the label on both sides
    noisy = ingest.load_cornell_dataset(str(tmp_path), name, feature_noise=0.5, feature_dim=10)
    assert noisy.x.shape == (mine.n_x, 10) and 0.3 < float((noisy.x - torch.nn.functional.pad(mine.x, (0, 10 - mine.x.shape[1]))).std()) < 0.7


@pytest.mark.skipif(not os.path.isfile(RAW_ZIP), reason='reference raw data not present')
def test_citation_loader_and_dataset_cache_equal_reference(tmp_path):
    """load_citation_dataset + the HypergraphDataset cache wrapper against the reference's loader on cora cocitation."""
    base = 'cocitation/cora/'
    _extract(tmp_path, [(base + f, 'raw/cora/' + f) for f in ('features.pickle', 'labels.pickle', 'hypergraph.pickle')])
    loaders, quiet = _reference_loaders()
    with quiet():
        ref = loaders.load_citation_dataset(path=str(tmp_path / 'raw'), dataset='cora')
    mine = ingest.load_citation_dataset(str(tmp_path / 'raw'), 'cora')
    assert mine.n_x == int(ref.n_x) and mine.num_hyperedges == int(ref.num_hyperedges)
    assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.y, ref.y) and torch.equal(mine.x, ref.x)
    ds = ingest.HypergraphDataset(root=str(tmp_path / 'pyg'), name='cora', p2raw=str(tmp_path / 'raw'))
    assert os.path.isfile(ds.processed_path) and ds.num_features == 1433 and ds.num_classes == 7
    assert int(ds.data.n_x[0]) == 2708 and torch.equal(ds.data.edge_index, ref.edge_index)
    again = ingest.HypergraphDataset(root=str(tmp_path / 'pyg'), name='cora', p2raw=None)       # served from the cache
    assert torch.equal(again.data.x, ref.x) and torch.equal(again.data.y, ref.y)
    with pytest.raises(ValueError):
        ingest.HypergraphDataset(root=str(tmp_path / 'pyg'), name='not-a-dataset')


@pytest.mark.skipif(not os.path.isfile(RAW_ZIP), reason='reference raw data not present')
def test_yelp_loader_equals_reference(tmp_path):
    files = ('yelp_restaurant_latlong.csv', 'yelp_restaurant_locations.csv', 'yelp_restaurant_name.csv',
             'yelp_restaurant_business_stars.csv', 'yelp_restaurant_incidence_H.csv')
    _extract(tmp_path, [('yelp/' + f, 'yelp/' + f) for f in files])
    loaders, quiet = _reference_loaders()
    with quiet():
        ref = loaders.load_yelp_dataset(path=str(tmp_path / 'yelp') + '/', dataset='yelp')
    mine = ingest.load_yelp_dataset(str(tmp_path / 'yelp'))
    assert mine.n_x == int(ref.n_x) == 50758 and mine.num_hyperedges == int(ref.num_hyperedges)
    assert torch.equal(mine.edge_index, ref.edge_index) and torch.equal(mine.y, ref.y)
    assert mine.x.shape == ref.x.shape and torch.equal(mine.x, ref.x)
