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
