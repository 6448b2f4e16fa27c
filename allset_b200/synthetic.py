"""Synthetic hypergraphs in the reference's incidence layout (BASELINE.json configs 3-5).

The reference hands `SetGNN` a COO list `edge_index [2, nnz]` with row 0 = node id ASCENDING and row 1 = hyperedge id
offset by N (reference src/preprocessing.py:394-409 `ExtractV2E` sorts by node; hyperedge ids start at N,
src/load_other_datasets.py:153-167).  These generators produce exactly that layout on any torch device so that the
CUDA path sees what it would see behind the reference's `train.py`: V->E targets arrive UNSORTED.

    poisson_hypergraph    hyperedge size = 1 + Poisson(mean_size - 1)                 (configs 3, 4)
    powerlaw_hypergraph   P(size = s) ~ s^-alpha on [lo, hi], one size-`hi` edge forced  (config 5)

Members are drawn uniformly WITH replacement (a repeated member inside one hyperedge has probability
~ size^2 / 2N, i.e. < 1e-4 per hyperedge at the configured sizes, and is legal input for the reference path, which
treats the list as a multiset).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def _assemble(sizes: torch.Tensor, n_nodes: int, gen: torch.Generator, sort_by_node: bool, he_offset: int
              ) -> torch.Tensor:
    dev = sizes.device
    m = sizes.numel()
    he = torch.repeat_interleave(torch.arange(m, device=dev, dtype=torch.int64), sizes)
    node = torch.randint(0, n_nodes, (he.numel(),), device=dev, dtype=torch.int64, generator=gen)
    if sort_by_node:
        node, order = torch.sort(node, stable=True)
        he = he[order]
    return torch.stack([node, he + he_offset])


def poisson_hypergraph(n_nodes: int, n_hyperedges: int, mean_size: float, seed: int = 1234,
                       device: Optional[torch.device] = None, sort_by_node: bool = True,
                       offset_hyperedge_ids: bool = True) -> torch.Tensor:
    """edge_index [2, nnz] int64; sizes 1 + Poisson(mean_size - 1)."""
    dev = torch.device(device if device is not None else 'cpu')
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    rate = torch.full((n_hyperedges,), float(mean_size) - 1.0, device=dev)
    sizes = 1 + torch.poisson(rate, generator=gen).long()
    return _assemble(sizes, n_nodes, gen, sort_by_node, n_nodes if offset_hyperedge_ids else 0)


def powerlaw_hypergraph(n_nodes: int, n_hyperedges: int, lo: int = 2, hi: int = 4096, alpha: float = 2.0,
                        seed: int = 1234, device: Optional[torch.device] = None, sort_by_node: bool = True,
                        offset_hyperedge_ids: bool = True) -> torch.Tensor:
    """edge_index [2, nnz] int64; P(size = s) proportional to s^-alpha for s in [lo, hi]; hyperedge 0 has size hi."""
    dev = torch.device(device if device is not None else 'cpu')
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    support = torch.arange(lo, hi + 1, device=dev, dtype=torch.float64)
    cdf = torch.cumsum(support.pow(-alpha), 0)
    cdf = cdf / cdf[-1]
    u = torch.rand(n_hyperedges, device=dev, dtype=torch.float64, generator=gen)
    sizes = lo + torch.searchsorted(cdf, u).clamp_(max=hi - lo)
    sizes[0] = hi
    return _assemble(sizes, n_nodes, gen, sort_by_node, n_nodes if offset_hyperedge_ids else 0)


def features(n_rows: int, d: int, dtype: torch.dtype = torch.bfloat16, seed: int = 1234,
             device: Optional[torch.device] = None, chunk_rows: int = 1 << 20) -> torch.Tensor:
    """X ~ N(0, 1) [n_rows, d] in `dtype`, generated chunk-wise so no fp32 copy of the whole matrix exists."""
    dev = torch.device(device if device is not None else 'cpu')
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    out = torch.empty((n_rows, d), dtype=dtype, device=dev)
    for r0 in range(0, n_rows, chunk_rows):
        r1 = min(n_rows, r0 + chunk_rows)
        out[r0:r1] = torch.randn((r1 - r0, d), device=dev, generator=gen).to(dtype)
    return out


def algorithmic_bytes(nnz: int, n_tgt: int, d: int, elem_bytes: int, heads: int = 0, weighted: bool = False,
                      stats: bool = False) -> int:
    """ALGORITHMIC bytes of one gather-reduce launch (gather model: no reuse credit):
    nnz * (d*s + 4) [row gather + int32 column id] + (n_tgt + 1) * 4 [rowptr] + n_tgt * d * s [output rows]
    + nnz * 4 if per-incidence weights are read; PMA (heads > 0) adds nnz * H * 4 for the gathered scores and
    n_tgt * H * 8 when the (max, sum) softmax statistics are written for backward."""
    b = nnz * (d * elem_bytes + 4) + (n_tgt + 1) * 4 + n_tgt * d * elem_bytes
    if weighted:
        b += nnz * 4
    if heads > 0:
        b += nnz * heads * 4
        if stats:
            b += n_tgt * heads * 8
    return int(b)
