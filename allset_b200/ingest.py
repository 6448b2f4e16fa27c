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
    key = np.unique(row0 * total + row1)                        # sort by (row 0, row 1) and drop duplicates
    ei = np.stack([key // total, key % total])
    return torch.from_numpy(ei), len(lists)


def _coalesced_star(nodes: np.ndarray, edges: np.ndarray, total: int) -> torch.Tensor:
    """[V|E ; E|V] sorted by (row 0, row 1) with duplicates dropped == torch_sparse.coalesce(edge_index, None, total, total)."""
    row0 = np.concatenate([nodes, edges]).astype(np.int64)
    row1 = np.concatenate([edges, nodes]).astype(np.int64)
    key = np.unique(row0 * total + row1)
    return torch.from_numpy(np.stack([key // total, key % total]))


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
