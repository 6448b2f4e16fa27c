"""Dataset ingest: the star-expansion incidence list in the reference's layout, and a flat binary
cache that replaces the reference's pickle -> `processed/data.pt` round trip (reference
src/convert_datasets_to_pygDataset.py:163-175, src/load_other_datasets.py:121-196).

    star_expansion(hyperedges, n_nodes)   == the edge_index built by load_citation_dataset (load_other_datasets.py:154-181):
                                             hyperedge ids n_nodes, n_nodes+1, ... in iteration order, the list
                                             [V|E ; E|V], sorted by (row 0, row 1), duplicates dropped (torch_sparse.coalesce)
    save_cache / load_cache               one file: a small JSON header + raw little-endian arrays, 64-byte aligned, read
                                          back with np.memmap (zero copy on the host, one H2D per array) -- the reference
                                          re-parses pickles into dense Python lists (`node_list += list(cur_he)`) on
                                          every first load and stores a pickled PyG `Data` afterwards

    load_le_dataset(dir, name)            == load_LE_dataset (load_other_datasets.py:32-119): `<name>.content` (id, features..., label
                                             per line, nodes first then hyperedges) + `<name>.edges` (node id, hyperedge id)
    load_cornell_dataset(dir, name)       == load_cornell_dataset (load_other_datasets.py:293-386): one hyperedge per line,
                                             one label per line; features = one-hot(label) + N(0, noise) (the reference
                                             draws the noise from numpy's global RNG, so features match in distribution,
                                             labels / incidence exactly)

    load_citation_dataset(dir, name)      == load_citation_dataset (load_other_datasets.py:121-196): the three HyperGCN pickles
    load_yelp_dataset(dir)                == load_yelp_dataset (load_other_datasets.py:198-291): five csv files
    HypergraphDataset(root, name, p2raw)  the `dataset_Hypergraph` wrapper train.py constructs (train.py:308-327) on top of
                                          those loaders and the binary cache

Host-side code (numpy / torch CPU); the arrays it yields are what `allset_b200.preprocessing` consumes on the GPU.
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Iterable, Mapping, Optional, Sequence, Union

import numpy as np
import torch

MAGIC = b'ALLSETB2'
_ALIGN = 64
_DTYPES = {'float32': np.float32, 'float16': np.float16, 'int64': np.int64, 'int32': np.int32, 'uint8': np.uint8,
           'bfloat16': np.uint16}


def star_expansion(hyperedges: Union[Mapping[object, Sequence[int]], Iterable[Sequence[int]]], n_nodes: int):
    """-> (edge_index [2, 2*nnz'] int64 in the reference layout, num_hyperedges).  `hyperedges`: dict name -> members
    (HyperGCN's hypergraph.pickle) or an iterable of member lists; ids are assigned in iteration order."""
    lists = list(hyperedges.values()) if isinstance(hyperedges, Mapping) else list(hyperedges)
    sizes = np.fromiter((len(h) for h in lists), dtype=np.int64, count=len(lists))
    nodes = np.fromiter((v for h in lists for v in h), dtype=np.int64, count=int(sizes.sum()))
    if nodes.size and (nodes.min() < 0 or nodes.max() >= n_nodes):
        raise ValueError('hyperedge member outside [0, %d)' % n_nodes)
    edges = np.repeat(np.arange(n_nodes, n_nodes + len(lists), dtype=np.int64), sizes)
    row0 = np.concatenate([nodes, edges])
    row1 = np.concatenate([edges, nodes])
    total = n_nodes + len(lists)                                # coalesce(m = n = edge_index.max() + 1)
    return _sorted_unique_pairs(row0, row1, total), len(lists)


def _coalesced_star(nodes: np.ndarray, edges: np.ndarray, total: int) -> torch.Tensor:
    """[V|E ; E|V] sorted by (row 0, row 1) with duplicates dropped == torch_sparse.coalesce(edge_index, None, total, total)."""
    row0 = np.concatenate([nodes, edges]).astype(np.int64)
    row1 = np.concatenate([edges, nodes]).astype(np.int64)
    return _sorted_unique_pairs(row0, row1, total)


def _sorted_unique_pairs(row0: np.ndarray, row1: np.ndarray, total: int) -> torch.Tensor:
    """Sort (row0, row1) pairs lexicographically and drop duplicates, through one int64 key per pair.  torch.unique
    (a radix sort) does 9 M keys in 0.3 s where numpy 2.3's np.unique takes 9 s on the yelp incidence list."""
    key = torch.unique(torch.from_numpy(row0 * total + row1), sorted=True)
    return torch.stack([torch.div(key, total, rounding_mode='floor'), key % total])


def load_le_dataset(path: str, dataset: str):
    """`<path>/<dataset>/<dataset>.content` + `.edges` -> namespace(x [n_x, F] f32, edge_index, y [n_x] i64, n_x,
    num_hyperedges), equal to the reference's load_LE_dataset (ids are remapped to their row position in .content; the
    reference does it with a Python dict + map over every incidence, here with one argsort + searchsorted)."""
    content = np.loadtxt(os.path.join(path, dataset, dataset + '.content'), dtype=np.float64, ndmin=2)
    ids = content[:, 0].astype(np.int64)
    labels = content[:, -1].astype(np.int64)
    feats = content[:, 1:-1].astype(np.float32)
    raw = np.loadtxt(os.path.join(path, dataset, dataset + '.edges'), dtype=np.int64, ndmin=2)
    order = np.argsort(ids, kind='stable')
    pos = np.searchsorted(ids[order], raw.reshape(-1))
    if np.any(pos >= ids.size) or np.any(ids[order][np.minimum(pos, ids.size - 1)] != raw.reshape(-1)):
        raise ValueError('%s.edges refers to ids that %s.content does not list' % (dataset, dataset))
    mapped = order[pos].reshape(raw.shape)
    nodes, edges = mapped[:, 0], mapped[:, 1]
    if nodes.max() != edges.min() - 1 or np.unique(mapped).size != mapped.max() + 1:
        raise ValueError('node / hyperedge ids must be consecutive with nodes first (reference asserts the same)')
    n_x = int(nodes.max()) + 1
    n_he = int(edges.max()) - n_x + 1
    return SimpleNamespace(x=torch.from_numpy(feats[:n_x].copy()), y=torch.from_numpy(labels[:n_x].copy()),
                           edge_index=_coalesced_star(nodes, edges, n_x + n_he), n_x=n_x, num_hyperedges=n_he)


def load_cornell_dataset(path: str, dataset: str, feature_noise: float = 0.1, feature_dim: Optional[int] = None,
                         generator: Optional[np.random.Generator] = None):
    """`node-labels-<dataset>.txt` + `hyperedges-<dataset>.txt` -> namespace(x, edge_index, y, n_x, num_hyperedges) as
    the reference's load_cornell_dataset: node ids shifted to start at 0, hyperedge ids n_x, n_x+1, ... in line order."""
    labels = np.loadtxt(os.path.join(path, dataset, 'node-labels-%s.txt' % dataset), dtype=np.int64, ndmin=1)
    n_x = labels.size
    n_cls = int(labels.max())
    feats = np.zeros((n_x, n_cls if feature_dim is None else max(feature_dim, n_cls)), dtype=np.float64)
    feats[np.arange(n_x), labels - 1] = 1
    rng = generator if generator is not None else np.random.default_rng()
    feats = rng.normal(feats, feature_noise, feats.shape)
    with open(os.path.join(path, dataset, 'hyperedges-%s.txt' % dataset)) as f:
        lines = [ln for ln in f.read().split('\n') if ln]
    sizes = np.fromiter((ln.count(',') + 1 for ln in lines), dtype=np.int64, count=len(lines))
    nodes = np.array(','.join(lines).split(','), dtype=np.int64)
    nodes = nodes - nodes.min()
    edges = np.repeat(np.arange(n_x, n_x + len(lines), dtype=np.int64), sizes)
    total = int(max(nodes.max(), edges.max())) + 1
    return SimpleNamespace(x=torch.from_numpy(feats.astype(np.float32)), y=torch.from_numpy(labels.copy()),
                           edge_index=_coalesced_star(nodes, edges, total), n_x=n_x, num_hyperedges=len(lines))


def load_citation_dataset(path: str, dataset: str):
    """`<path>/<dataset>/{features,labels,hypergraph}.pickle` (HyperGCN's cocitation / coauthorship layout) -> namespace(x,
    edge_index, y, n_x, num_hyperedges) equal to the reference's load_citation_dataset (load_other_datasets.py:121-196):
    dense float features, the star expansion of the hyperedge dictionary, labels as given."""
    import pickle
    d = os.path.join(path, dataset)
    with open(os.path.join(d, 'features.pickle'), 'rb') as f:
        feats = pickle.load(f)
    with open(os.path.join(d, 'labels.pickle'), 'rb') as f:
        labels = pickle.load(f)
    with open(os.path.join(d, 'hypergraph.pickle'), 'rb') as f:
        hyperedges = pickle.load(f)
    x = np.asarray(feats.todense() if hasattr(feats, 'todense') else feats, dtype=np.float32)
    y = torch.as_tensor(np.asarray(labels), dtype=torch.int64)
    if x.shape[0] != y.numel():
        raise ValueError('%d feature rows but %d labels' % (x.shape[0], y.numel()))
    edge_index, n_he = star_expansion(hyperedges, x.shape[0])
    return SimpleNamespace(x=torch.from_numpy(x), y=y, edge_index=edge_index, n_x=int(x.shape[0]), num_hyperedges=n_he)


def load_yelp_dataset(path: str, dataset: str = 'yelp', name_dictionary_size: int = 1000):
    """The yelp restaurant hypergraph (reference load_other_datasets.py:198-291): nodes = restaurants, one hyperedge per
    user.  Features = [latitude, longitude | one-hot state | one-hot city | bag of words of the name (the 1000 most
    frequent terms; sklearn CountVectorizer with the reference's settings)], labels = star bins, incidence from the
    (node, he) pairs of `yelp_restaurant_incidence_H.csv` (1-based in the file).  `path` holds the five csv files."""
    import pandas as pd
    from sklearn.feature_extraction.text import CountVectorizer
    latlong = pd.read_csv(os.path.join(path, 'yelp_restaurant_latlong.csv')).values
    loc = pd.read_csv(os.path.join(path, 'yelp_restaurant_locations.csv'))
    n_x = loc.shape[0]
    rows = np.arange(n_x)

    def one_hot(ids):
        m = np.zeros((n_x, int(ids.max())), dtype=np.float64)
        m[rows, ids - 1] = 1
        return m

    names = pd.read_csv(os.path.join(path, 'yelp_restaurant_name.csv')).values.reshape(-1)
    bow = CountVectorizer(max_features=name_dictionary_size, stop_words='english', strip_accents='ascii').fit_transform(names)
    feats = np.hstack([latlong, one_hot(loc.state_int.values), one_hot(loc.city_int.values), np.asarray(bow.todense())])
    labels = pd.read_csv(os.path.join(path, 'yelp_restaurant_business_stars.csv')).values.reshape(-1)
    if labels.size != n_x or feats.shape[0] != n_x:
        raise ValueError('yelp: %d locations, %d labels, %d feature rows' % (n_x, labels.size, feats.shape[0]))
    H = pd.read_csv(os.path.join(path, 'yelp_restaurant_incidence_H.csv'))
    nodes = H.node.values.astype(np.int64) - 1
    edges = H.he.values.astype(np.int64) - 1 + n_x
    total = int(max(nodes.max(), edges.max())) + 1
    return SimpleNamespace(x=torch.from_numpy(feats.astype(np.float32)), y=torch.from_numpy(labels.astype(np.int64)),
                           edge_index=_coalesced_star(nodes, edges, total), n_x=n_x, num_hyperedges=int(H.he.values.max()))


# dataset name -> (directory under the raw root, loader); the names reference train.py accepts (train.py:293-299)
_LE = ('20newsW100', 'ModelNet40', 'zoo', 'NTU2012', 'Mushroom')
_CORNELL = ('amazon-reviews', 'walmart-trips', 'house-committees')


class HypergraphDataset(object):
    """Replacement for the reference's `dataset_Hypergraph` (convert_datasets_to_pygDataset.py:39-175): same constructor
    arguments and the attributes train.py reads (`.data`, `.num_features`, `.num_classes`), but the first load parses the
    raw files with the vectorised loaders above and writes ONE flat binary cache (`<root>/<name>/processed/data[...].allset`)
    that later loads memory-map -- instead of pickling a list-built PyG `Data` twice (raw/ and processed/data.pt)."""

    def __init__(self, root='../data/pyg_data/hypergraph_dataset_updated/', name=None, p2raw=None, train_percent=0.01,
                 feature_noise=None, transform=None, pre_transform=None):
        known = _LE + _CORNELL + ('walmart-trips-100', 'house-committees-100', 'coauthor_cora', 'coauthor_dblp', 'yelp',
                                  'cora', 'citeseer', 'pubmed')
        if name not in known:
            raise ValueError('name of hypergraph dataset must be one of: %s' % (list(known),))
        if p2raw is not None and not os.path.isdir(p2raw):
            raise ValueError('path to raw hypergraph dataset "%s" does not exist!' % p2raw)
        self.name, self.root, self.p2raw, self.feature_noise = name, root, p2raw, feature_noise
        self._train_percent = train_percent
        fname = 'data.allset' if feature_noise is None else 'data_noise_%s.allset' % feature_noise
        self.processed_path = os.path.join(root, name, 'processed', fname)
        if not os.path.isfile(self.processed_path):
            os.makedirs(os.path.dirname(self.processed_path), exist_ok=True)
            d = self._parse()
            if pre_transform is not None:
                d = pre_transform(d)
            save_cache(self.processed_path, d.x, d.edge_index, d.y, n_x=d.n_x, num_hyperedges=d.num_hyperedges)
        self.data = load_cache(self.processed_path)
        # train.py indexes these as 1-element tensors (train.py:334-339; PyG's collate made them so)
        self.data.n_x = torch.tensor([self.data.n_x])
        self.data.num_hyperedges = torch.tensor([self.data.num_hyperedges])
        self.data.train_percent = torch.tensor([train_percent])
        self.train_percent = train_percent
        if transform is not None:
            self.data = transform(self.data)

    def _parse(self):
        n, raw = self.name, self.p2raw
        if raw is None:
            raise ValueError('no cache at %s and no p2raw to build it from' % self.processed_path)
        if n in ('cora', 'citeseer', 'pubmed'):
            return load_citation_dataset(raw, n)
        if n in ('coauthor_cora', 'coauthor_dblp'):
            return load_citation_dataset(raw, n.split('_')[-1])
        if n in _CORNELL or n in ('walmart-trips-100', 'house-committees-100'):
            if self.feature_noise is None:
                raise ValueError('for cornell datasets, feature noise cannot be %s' % self.feature_noise)
            dim = int(n.split('-')[-1]) if n.endswith('-100') else None
            base = '-'.join(n.split('-')[:-1]) if n.endswith('-100') else n
            return load_cornell_dataset(raw, base, feature_noise=float(self.feature_noise), feature_dim=dim)
        if n == 'yelp':
            return load_yelp_dataset(raw, n)
        return load_le_dataset(raw, n)

    @property
    def num_features(self) -> int:
        return int(self.data.x.shape[1])

    @property
    def num_classes(self) -> int:
        y = self.data.y
        return int(y.max()) + 1 if y.dim() == 1 else int(y.shape[1])

    def __repr__(self):
        return '{}()'.format(self.name)


def _to_numpy(t: torch.Tensor):
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.uint16).numpy(), 'bfloat16'
    a = t.numpy()
    if a.dtype.name not in _DTYPES:
        raise TypeError('unsupported dtype %s' % a.dtype)
    return a, a.dtype.name


def save_cache(path: str, x: torch.Tensor, edge_index: torch.Tensor, y: Optional[torch.Tensor] = None, *,
               n_x: int, num_hyperedges: int, **extra: torch.Tensor) -> None:
    """Write one cache file: MAGIC, u64 header length, JSON header, then every array at a 64-byte aligned offset."""
    arrays = {'x': x, 'edge_index': edge_index}
    if y is not None:
        arrays['y'] = y
    arrays.update(extra)
    meta, blobs, off = {}, [], 0
    for name, t in arrays.items():
        a, dt = _to_numpy(t)
        off = (off + _ALIGN - 1) // _ALIGN * _ALIGN
        meta[name] = {'dtype': dt, 'shape': list(a.shape), 'offset': off, 'nbytes': int(a.nbytes)}
        blobs.append((off, a))
        off += a.nbytes
    header = json.dumps({'version': 1, 'n_x': int(n_x), 'num_hyperedges': int(num_hyperedges), 'arrays': meta}).encode()
    base = (len(MAGIC) + 8 + len(header) + _ALIGN - 1) // _ALIGN * _ALIGN
    tmp = path + '.tmp'
    with open(tmp, 'wb') as f:
        f.write(MAGIC)
        f.write(np.uint64(len(header)).tobytes())
        f.write(header)
        for o, a in blobs:
            f.seek(base + o)
            f.write(a.tobytes(order='C'))
        f.truncate(base + off)
    os.replace(tmp, path)                                       # readers never see a partial file


def load_cache(path: str, device: Union[str, torch.device, None] = None, pin: bool = False):
    """-> namespace(x, edge_index, y?, n_x, num_hyperedges, ...) with the attributes reference train.py reads off `data`.
    Arrays are memory-mapped; `device` moves them (one copy each), `pin` stages them in pinned host memory first."""
    with open(path, 'rb') as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError('%s is not an allset_b200 cache file' % path)
        hlen = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
        header = json.loads(f.read(hlen).decode())
    if header.get('version') != 1:
        raise ValueError('unsupported cache version %r' % header.get('version'))
    base = (len(MAGIC) + 8 + hlen + _ALIGN - 1) // _ALIGN * _ALIGN
    out = SimpleNamespace(n_x=header['n_x'], num_hyperedges=header['num_hyperedges'])
    for name, m in header['arrays'].items():
        # copy-on-write mapping: pages are read lazily and the tensor is writable without touching the file
        a = np.memmap(path, mode='c', dtype=_DTYPES[m['dtype']], offset=base + m['offset'], shape=tuple(m['shape']))
        t = torch.from_numpy(a)
        if m['dtype'] == 'bfloat16':
            t = t.view(torch.bfloat16)
        if pin:
            t = t.pin_memory()
        if device is not None:
            t = t.to(device, non_blocking=pin)
        setattr(out, name, t)
    return out
