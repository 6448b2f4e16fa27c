"""Vectorised, device-agnostic restatement of the reference's incidence preprocessing: the step that
produces the exact `data.edge_index` / `data.norm` layout the aggregation path consumes.

The reference does this on the CPU with O(N) Python loops (`for i in range(num_nodes): if i not in skip_node_lst`,
a `Counter` over every incidence -- reference src/preprocessing.py:423-441), which takes minutes at the 10M-vertex
configs.  Here each function is a handful of sort / bincount / mask operations that run on whatever device the
incidence list lives on (the GPU in practice).  Results equal the reference's (tests/test_preprocessing.py, fixtures
recorded from the reference's own functions) up to the order of one node's hyperedges, which the reference leaves to
an UNSTABLE `torch.sort` (preprocessing.py:398,446); here the sort is stable, i.e. original order within a node.

    extract_v2e        == ExtractV2E        (preprocessing.py:394-409)
    add_self_loops     == Add_Self_Loops    (preprocessing.py:412-448)
    norm_construction  == norm_contruction  (preprocessing.py:451-464, TYPE='V2E')
    expand_edge_index  == expand_edge_index (preprocessing.py:22-144, the --exclude_self option of train.py:348-349)
    preprocess         == the train.py:344-353 sequence for AllDeepSets / AllSetTransformer
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def extract_v2e(edge_index: torch.Tensor, n_nodes: int, n_hyperedges: int) -> torch.Tensor:
    """Keep the V->E half of a star-expansion list [V|E ; E|V]: columns whose row 0 is a node id, sorted by node.
    Raises where the reference prints 'num_hyperedges does not match! 1' and returns None."""
    if edge_index.numel() == 0 or int(edge_index[0].max()) != n_nodes + n_hyperedges - 1:
        raise ValueError('num_hyperedges does not match! 1')
    order = torch.sort(edge_index[0], stable=True)[1]
    ei = edge_index[:, order].long()
    keep = int((ei[0] < n_nodes).sum())                  # == first position where row 0 == n_nodes
    return ei[:, :keep].contiguous()


def add_self_loops(edge_index: torch.Tensor, n_nodes: int, n_hyperedges: int) -> Tuple[torch.Tensor, int]:
    """One new size-1 hyperedge per node that is not already the sole member of a size-1 hyperedge; new ids continue
    from edge_index[1].max()+1 in ascending node order; result re-sorted by node.  Returns (edge_index, totedges)."""
    if edge_index.numel() == 0 or int(edge_index[1].max()) != n_nodes + n_hyperedges - 1:
        raise ValueError('num_hyperedges does not match! 2')
    node, he = edge_index[0], edge_index[1]
    base = int(he.min())
    size = torch.bincount(he - base)
    in_singleton = size[he - base] == 1                   # incidences that are a whole hyperedge on their own
    has_loop = torch.zeros(n_nodes, dtype=torch.bool, device=edge_index.device)
    has_loop[node[in_singleton]] = True
    new_nodes = torch.nonzero(~has_loop, as_tuple=False).flatten()
    first_new = int(he.max()) + 1
    new_ids = first_new + torch.arange(new_nodes.numel(), device=edge_index.device, dtype=edge_index.dtype)
    ei = torch.cat([edge_index, torch.stack([new_nodes.to(edge_index.dtype), new_ids])], dim=1)
    order = torch.sort(ei[0], stable=True)[1]
    return ei[:, order].long().contiguous(), int(n_hyperedges + new_nodes.numel())


def norm_construction(edge_index: torch.Tensor, option: str = 'all_one') -> torch.Tensor:
    """Per-incidence weights `data.norm`: 'all_one' -> int64 ones (the reference's dtype, preprocessing.py:454);
    'deg_half_sym' -> deg(v)^-1/2 * |e|^-1/2 in float32."""
    if option == 'all_one':
        return torch.ones_like(edge_index[0])
    if option == 'deg_half_sym':
        node, he = edge_index[0], edge_index[1] - edge_index[1].min()
        v_deg = torch.bincount(node)
        e_deg = torch.bincount(he)
        return v_deg.pow(-0.5)[node] * e_deg.pow(-0.5)[he]
    raise ValueError("normtype must be 'all_one' or 'deg_half_sym', got %r" % (option,))


def expand_edge_index(edge_index: torch.Tensor, n_nodes: int, n_hyperedges: int, edge_th: int = 0) -> torch.Tensor:
    """"Exclude self" expansion (reference preprocessing.py:22-144): a hyperedge e = {n_1..n_s}, s > 1, becomes s new
    hyperedges e_1..e_s (one per member, numbered in member order), and node n_j joins every e_i with i != j -- so that
    the message a node receives from "its" hyperedge excludes its own features.  Size-1 hyperedges are kept as one new
    id; with edge_th > 0 hyperedges larger than edge_th are dropped.  New ids start at n_nodes and are handed out in
    ascending original-hyperedge order; the result is sorted by node (stable).  s(s-1) incidences per hyperedge.

    The reference does this with a Python loop over every hyperedge and a boolean mask over the WHOLE list per
    hyperedge (O(M * nnz)); here it is one stable sort plus repeat_interleave / arithmetic on the device.
    n_hyperedges = data.totedges after Add_Self_Loops, else data.num_hyperedges (ids [n_nodes, n_nodes + n_hyperedges))."""
    dev = edge_index.device
    node, he = edge_index[0].long(), edge_index[1].long() - n_nodes
    inside = (he >= 0) & (he < n_hyperedges)                   # the reference only visits ids in that range
    node, he = node[inside], he[inside]
    order = torch.sort(he, stable=True)[1]                     # member order inside a hyperedge = list order (:63)
    node, he = node[order], he[order]
    size = torch.bincount(he, minlength=n_hyperedges)
    keep_e = size > 0
    if edge_th > 0:
        keep_e &= size <= edge_th
    new_ids = torch.where(keep_e, size, torch.zeros_like(size))          # ids handed out per hyperedge (1 if s == 1)
    base = n_nodes + torch.cumsum(new_ids, 0) - new_ids                   # first new id of each hyperedge
    start = torch.cumsum(size, 0) - size                                  # first list position of each hyperedge
    pos = torch.arange(he.numel(), device=dev) - start[he]                # member index j of every incidence
    s_inc = torch.where(keep_e[he], size[he], torch.zeros_like(he))       # size of the incidence's hyperedge (0 = dropped)
    single = s_inc == 1
    # members of size-1 hyperedges: one incidence (n, base_e)
    out_node = [node[single]]
    out_he = [base[he[single]]]
    # members of larger hyperedges: node n_j x every new id base_e + i, i != j
    multi = s_inc > 1
    nj, ej, jj, sj = node[multi], he[multi], pos[multi], s_inc[multi]
    rep_node = torch.repeat_interleave(nj, sj)
    rep_e = torch.repeat_interleave(ej, sj)
    rep_j = torch.repeat_interleave(jj, sj)
    first = torch.cumsum(sj, 0) - sj
    ii = torch.arange(rep_node.numel(), device=dev) - torch.repeat_interleave(first, sj)      # i = 0..s-1
    keep = ii != rep_j
    out_node.append(rep_node[keep])
    out_he.append(base[rep_e[keep]] + ii[keep])
    ei = torch.stack([torch.cat(out_node), torch.cat(out_he)])
    order = torch.sort(ei[0], stable=True)[1]
    return ei[:, order].contiguous()


def preprocess(edge_index: torch.Tensor, n_nodes: int, n_hyperedges: int, add_self_loop: bool = True,
               normtype: str = 'all_one', star_expansion: bool = True, exclude_self: bool = False):
    """train.py:344-353 for method in {AllSetTransformer, AllDeepSets}: ExtractV2E -> Add_Self_Loops ->
    [expand_edge_index] -> norm.
    Returns (edge_index [2, nnz] int64, norm [nnz], total hyperedges)."""
    ei = extract_v2e(edge_index, n_nodes, n_hyperedges) if star_expansion else edge_index
    tot = n_hyperedges
    if add_self_loop:
        ei, tot = add_self_loops(ei, n_nodes, n_hyperedges)
    if exclude_self:                                          # train.py:348-349
        ei = expand_edge_index(ei, n_nodes, tot)
        tot = int(ei[1].max()) - n_nodes + 1 if ei.numel() else 0
    return ei, norm_construction(ei, normtype), tot
