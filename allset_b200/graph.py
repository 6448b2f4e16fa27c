"""Incidence container: the bipartite vertex-hyperedge COO list, sorted once into two int32 CSRs.

The reference keeps `data.edge_index [2, nnz]` (row 0 = node, row 1 = hyperedge; reference
src/preprocessing.py:394-447) and lets torch_scatter rediscover the output size (`index.max()+1`, a D2H sync)
and fight over unsorted targets with atomics on EVERY call (reference src/layers.py:656, :194).  Here the list is
sorted once per graph (stable, so per-segment order = COO order) into

    by_tgt : rowptr[n_tgt+1], col[nnz] = source row,  perm[nnz] = COO position     (forward gather)
    by_src : rowptr[n_src+1], col[nnz] = target row,  perm[nnz]                    (backward = transposed gather)

and cached on the tensor object itself together with its `_version`, because the reference mutates `edge_index`
in place (src/models.py:453-454).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

LONG_SEGMENT_THRESHOLD = 1024     # segments longer than this get a CTA instead of a lane group


class Csr(object):
    """CSR by target row of one direction of the incidence list (all int32, on the graph's device)."""

    def __init__(self, rowptr, col, perm, n_tgt: int, n_src: int, threshold: int = LONG_SEGMENT_THRESHOLD):
        self.rowptr, self.col, self.perm = rowptr, col, perm
        self.n_tgt, self.n_src, self.nnz = int(n_tgt), int(n_src), int(col.numel())
        self.long_threshold = int(threshold)
        self.long_ids = _lib.long_segments(rowptr, self.n_tgt, self.long_threshold) if self.nnz > threshold else None
        # longest segment (one sync at graph build): moderate lengths stay inside the stream kernels
        self.max_len = int((rowptr[1:self.n_tgt + 1] - rowptr[:self.n_tgt]).max()) if (self.long_ids is not None and self.n_tgt > 0) else 0
        self._perm64 = None
        self._inv_count = None

    @property
    def perm64(self) -> torch.Tensor:
        """perm as int64 (torch indexing wants long)."""
        if self._perm64 is None:
            self._perm64 = self.perm.long()
        return self._perm64

    @property
    def inv_count(self) -> torch.Tensor:
        """1 / max(segment length, 1) per target row (torch_scatter's scatter_mean clamps the count to 1)."""
        if self._inv_count is None:
            cnt = (self.rowptr[1:self.n_tgt + 1] - self.rowptr[:self.n_tgt]).clamp_(min=1)
            self._inv_count = 1.0 / cnt.float()
        return self._inv_count

    def head(self, n_tgt: int) -> 'Csr':
        """The same CSR restricted to its first n_tgt target rows (all incidences must live there)."""
        if n_tgt == self.n_tgt:
            return self
        if n_tgt > self.n_tgt:
            raise ValueError('cannot grow a CSR')
        c = Csr.__new__(Csr)
        c.__dict__.update(self.__dict__)
        c.n_tgt = int(n_tgt)
        c._inv_count = None
        if c.long_ids is not None:
            c.long_ids = c.long_ids[c.long_ids < n_tgt].contiguous()
            if c.long_ids.numel() == 0:
                c.long_ids = None
        return c


class Incidence(object):
    """One direction of message passing: rows of `src` (gathered) -> rows of `tgt` (reduced).

    `by_tgt` drives the forward kernels, `by_src` their gradients.  `reversed()` swaps the roles without
    re-sorting (V->E and E->V share the two CSRs)."""

    def __init__(self, by_tgt: Csr, by_src: Csr):
        self.by_tgt, self.by_src = by_tgt, by_src
        self._resized = {}

    @property
    def n_src(self) -> int:
        return self.by_src.n_tgt

    @property
    def n_tgt(self) -> int:
        return self.by_tgt.n_tgt

    @property
    def nnz(self) -> int:
        return self.by_tgt.nnz

    @classmethod
    def from_coo(cls, src: torch.Tensor, tgt: torch.Tensor, n_src: Optional[int] = None,
                 n_tgt: Optional[int] = None, validate: bool = True) -> 'Incidence':
        """src/tgt: int64 [nnz] CUDA tensors.  n_tgt defaults to tgt.max()+1 (torch_scatter's implicit size rule,
        reference src/layers.py:641-656 drops dim_size on purpose); n_src defaults to src.max()+1."""
        if not src.is_cuda or not tgt.is_cuda:
            raise RuntimeError('allset_b200 has no CPU path: the incidence list must live on a CUDA device')
        if src.dim() != 1 or src.shape != tgt.shape:
            raise ValueError('src and tgt must be 1-D tensors of equal length')
        src = src.long().contiguous()
        tgt = tgt.long().contiguous()
        nnz = src.numel()
        if nnz > 0 and (validate or n_src is None or n_tgt is None):
            lo_s, hi_s = torch.aminmax(src)
            lo_t, hi_t = torch.aminmax(tgt)
            lo_s, hi_s, lo_t, hi_t = int(lo_s), int(hi_s), int(lo_t), int(hi_t)    # one sync, at graph build only
            if lo_s < 0 or lo_t < 0:
                raise IndexError('negative row id in the incidence list')
            n_src = hi_s + 1 if n_src is None else n_src
            n_tgt = hi_t + 1 if n_tgt is None else n_tgt
            if hi_s >= n_src or hi_t >= n_tgt:
                raise IndexError('incidence list refers to row %d / %d outside [0,%d) / [0,%d)'
                                 % (hi_s, hi_t, n_src, n_tgt))
        n_src = 0 if n_src is None else int(n_src)
        n_tgt = 0 if n_tgt is None else int(n_tgt)
        rp_t, col_t, perm_t = _lib.csr_from_coo(tgt, src, n_tgt)
        rp_s, col_s, perm_s = _lib.csr_from_coo(src, tgt, n_src)
        return cls(Csr(rp_t, col_t, perm_t, n_tgt, n_src), Csr(rp_s, col_s, perm_s, n_src, n_tgt))

    def reversed(self, n_tgt: Optional[int] = None) -> 'Incidence':
        """Swap source and target.  n_tgt < n_src of this object reproduces torch_scatter dropping trailing rows
        that receive nothing (E->V output has max(node)+1 rows)."""
        by_tgt = self.by_src if n_tgt is None else self.by_src.head(n_tgt)
        return Incidence(by_tgt, self.by_tgt)

    def with_n_src(self, n_src: int) -> 'Incidence':
        """The same incidence list over a source table of `n_src` rows (>= max(src)+1).  SetGNN needs it when trailing
        nodes belong to no hyperedge: the E->V output then has max(node)+1 < N rows (torch_scatter's implicit size,
        SURVEY Appendix A) and feeds the next layer's V->E, which the reference -- a plain gather -- accepts."""
        if n_src == self.n_src:
            return self
        hit = self._resized.get(n_src)
        if hit is None:
            if n_src > self.n_src:
                raise ValueError('x has %d rows but the incidence list was built for %d source rows' % (n_src, self.n_src))
            s = self.by_src
            if n_src < self.n_src and int(s.rowptr[n_src]) != s.nnz:
                raise IndexError('incidence list refers to source rows beyond the %d rows of x' % n_src)
            t = Csr.__new__(Csr)
            t.__dict__.update(self.by_tgt.__dict__)
            t.n_src = int(n_src)
            hit = self._resized[n_src] = Incidence(t, s.head(n_src))
        return hit

    def weights_all_one(self, norm: torch.Tensor) -> bool:
        """True when the per-incidence weights are all exactly 1 so the multiply can be skipped.  Only INTEGER weights
        take the shortcut -- the reference default `data.norm = ones_like(edge_index[0])` is int64
        (src/preprocessing.py:453-454); float weights (deg_half_sym, `Importance * norm`) always go through the weighted
        kernel.  The verdict is cached ON the tensor object with its `_version` (one sync the first time): a cache keyed
        on the storage address would be fooled by the allocator handing the same address to a later temporary."""
        if norm.is_floating_point() or norm.is_complex():
            return False
        tagged = getattr(norm, _ONES_ATTR, None)
        if tagged is not None and tagged[0] == norm._version:
            return tagged[1]
        hit = bool((norm == 1).all().item()) if norm.numel() > 0 else True
        try:
            setattr(norm, _ONES_ATTR, (norm._version, hit))
        except Exception:  # noqa  (tensor subclasses without a __dict__)
            pass
        return hit


_ONES_ATTR = '_allset_all_one'

# ------------------------------------------------------------------------------------------------------------
# lookup for layers called with a plain `edge_index` tensor
# ------------------------------------------------------------------------------------------------------------
_ATTR = '_allset_incidence'


def attach(edge_index: torch.Tensor, inc: Incidence) -> torch.Tensor:
    """Remember `inc` on the tensor object itself (dies with it; invalidated by in-place edits via _version)."""
    setattr(edge_index, _ATTR, (edge_index._version, inc))
    return edge_index


def incidence_of(edge_index: torch.Tensor, n_src: int) -> Incidence:
    """Incidence for `edge_index [2, nnz]` (row 0 = source row ids, row 1 = target row ids, the PyG
    flow='source_to_target' convention the reference layers use).  Built on first sight of a tensor object and
    cached on it, so a training loop that reuses `data.edge_index` sorts once."""
    tagged = getattr(edge_index, _ATTR, None)
    if tagged is not None and tagged[0] == edge_index._version and tagged[1].n_src == n_src:
        return tagged[1]
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError('edge_index must be [2, nnz]')
    inc = Incidence.from_coo(edge_index[0], edge_index[1], n_src=n_src)
    attach(edge_index, inc)
    return inc
