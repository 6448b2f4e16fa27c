"""Multi-GPU partition of the V->E / E->V path: one process per GPU, target segments split into contiguous ranges.

The reference is single-device; this is the B200 scale-out named by BASELINE.json:north_star.
Each direction shards naturally BY TARGET SEGMENT: hyperedges are independent units of V->E given
read access to the replicated vertex rows, vertices are independent units of E->V given the hyperedge rows.  So

    V->E : rank r reduces hyperedges [e_lo, e_hi) from the replicated X_v           (no collective inside)
    exch : all-gather of X_e rows   (small: |E| d s bytes)
    E->V : rank r reduces vertices   [v_lo, v_hi) from the gathered X_e             (no collective inside)
    exch : all-gather of the updated X_v rows -- THE all-gather of north_star, |V| d s bytes over NVLink

Ranges are contiguous in the CSR-by-target, so a shard is a slice of (rowptr, col): no re-sort, no copy of `col`.
Boundaries balance the number of INCIDENCES (bytes gathered), not the number of rows, which matters for the
power-law config.  Pure index arithmetic: works on any device (the gloo tests run it on CPU tensors).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def balanced_ranges(rowptr: torch.Tensor, parts: int, row_weight: int = 0) -> List[Tuple[int, int]]:
    """Split rows [0, n) of a CSR into `parts` contiguous ranges of (nearly) equal cost, where the cost of row t is
    (rowptr[t+1] - rowptr[t]) + row_weight  -- incidences gathered plus `row_weight` incidence-equivalents for
    writing the output row.  Returns [(lo, hi)] * parts covering [0, n) in order; ranges may be empty."""
    n = int(rowptr.numel()) - 1
    if parts <= 0:
        raise ValueError('parts must be positive')
    if n <= 0:
        return [(0, 0)] * parts
    cost = rowptr.to(torch.int64) + row_weight * torch.arange(n + 1, device=rowptr.device, dtype=torch.int64)
    total = int(cost[-1])
    targets = torch.tensor([(total * k) // parts for k in range(1, parts)], device=rowptr.device, dtype=torch.int64)
    cuts = torch.searchsorted(cost, targets, right=False).clamp_(max=n).tolist() if parts > 1 else []
    bounds = [0] + cuts + [n]
    for i in range(1, len(bounds)):                        # monotone (searchsorted already is; guard the clamp)
        bounds[i] = max(bounds[i], bounds[i - 1])
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]


def slice_csr(rowptr: torch.Tensor, col: torch.Tensor, lo: int, hi: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Rows [lo, hi) of a CSR as (rowptr_local [hi-lo+1] starting at 0, col_local view, first incidence offset)."""
    if not (0 <= lo <= hi <= rowptr.numel() - 1):
        raise ValueError('bad range [%d, %d) for %d rows' % (lo, hi, rowptr.numel() - 1))
    p0, p1 = int(rowptr[lo]), int(rowptr[hi])
    return (rowptr[lo:hi + 1] - rowptr[lo]).contiguous(), col[p0:p1], p0


def row_views(full: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> List[torch.Tensor]:
    """Views of the replicated [rows, d] buffer, one per rank's range -- the receive list of the uneven all-gather."""
    return [full[lo:hi] for lo, hi in ranges]


def allgather_rows(full: torch.Tensor, ranges: Sequence[Tuple[int, int]], rank: int, group=None, async_op=False):
    """All-gather row ranges IN PLACE into the replicated buffer: rank r owns full[ranges[r]] (already written by its
    kernel) and receives every other range.  Equal-sized ranges take the single-kernel all_gather_into_tensor path;
    ragged ones one broadcast per non-empty range (what NCCL's uneven all-gather decomposes into; also the only
    form gloo accepts).  With async_op the returned handle(s) must be waited on by the caller."""
    import torch.distributed as dist
    sizes = {hi - lo for lo, hi in ranges}
    mine = full[ranges[rank][0]:ranges[rank][1]]
    contiguous_cover = all(ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))
    if len(sizes) == 1 and contiguous_cover and ranges[0][0] == 0:
        rows = ranges[-1][1]
        return dist.all_gather_into_tensor(full[:rows], mine, group=group, async_op=async_op)
    works = []
    for r, view in enumerate(row_views(full, ranges)):
        if view.shape[0] > 0:
            src = r if group is None else dist.get_global_rank(group, r)
            works.append(dist.broadcast(view, src=src, group=group, async_op=async_op))
    return works if async_op else None


def equal_ranges(n: int, parts: int) -> List[Tuple[int, int]]:
    """ceil(n / parts) rows per rank (the last ranges may be shorter or empty)."""
    per = (n + parts - 1) // parts
    return [(min(n, r * per), min(n, (r + 1) * per)) for r in range(parts)]


def choose_ranges(rowptr: torch.Tensor, parts: int, row_weight: int = 0, tolerance: float = 0.02
                  ) -> List[Tuple[int, int]]:
    """Equal row counts when that is also cost-balanced to within `tolerance` (random graphs: law of large numbers;
    it keeps the exchange on the one-kernel all_gather_into_tensor path), cost-balanced ragged ranges otherwise."""
    n = int(rowptr.numel()) - 1
    if parts == 1 or n == 0:
        return [(0, n)] + [(n, n)] * (parts - 1)
    if n % parts == 0:
        eq = equal_ranges(n, parts)
        b = torch.tensor([lo for lo, _ in eq] + [n], device=rowptr.device)
        cost = (rowptr[b].to(torch.int64) + row_weight * b.to(torch.int64)).tolist()
        per = [cost[i + 1] - cost[i] for i in range(parts)]
        mean = sum(per) / parts
        if mean == 0 or max(per) <= (1.0 + tolerance) * mean:
            return eq
    return balanced_ranges(rowptr, parts, row_weight)


def range_shapes(rowptr: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """(rows, longest segment) of every range of a CSR -- one host sync, at partition time."""
    n = int(rowptr.numel()) - 1
    if n <= 0:
        return [(0, 0) for _ in ranges]
    lens = (rowptr[1:] - rowptr[:-1])
    longest = torch.stack([lens[lo:hi].max() if hi > lo else lens.new_zeros(()) for lo, hi in ranges]).tolist()
    return [(hi - lo, int(m)) for (lo, hi), m in zip(ranges, longest)]


def peer_need_mask(rowptr: torch.Tensor, col: torch.Tensor, rows: int, ranges: Sequence[Tuple[int, int]], rank: int
                   ) -> torch.Tensor:
    """For a CSR slice (row r lists the OTHER side's ids it is incident to) and the other side's partition `ranges`:
    uint8 [rows], bit j set iff row r has an incidence inside the range of the j-th other rank (ascending, skipping
    `rank`) -- i.e. that rank will gather row r.  Pure index arithmetic (any device)."""
    dev = rowptr.device
    mask = torch.zeros(rows, dtype=torch.uint8, device=dev)
    nnz = int(col.numel())
    if rows == 0 or nnz == 0:
        return mask
    lens = (rowptr[1:rows + 1] - rowptr[:rows]).long()
    row_of = torch.repeat_interleave(torch.arange(rows, device=dev), lens)
    bounds = torch.tensor([hi for _, hi in ranges], device=dev, dtype=col.dtype)
    owner = torch.bucketize(col[:nnz], bounds, right=True)             # range index of every incidence
    j = 0
    for q in range(len(ranges)):
        if q == rank:
            continue
        hit = torch.zeros(rows, dtype=torch.bool, device=dev)
        hit[row_of[owner == q]] = True
        mask |= hit.to(torch.uint8) << j
        j += 1
    return mask


class ReplicatedRows(object):
    """A [rows, d] feature buffer replicated on every rank of the group, allocated in SYMMETRIC memory
    (torch.distributed._symmetric_memory) so that each rank holds peer-mapped pointers to all replicas: the fused
    kernels store the rows they reduce straight into every replica over NVLink.  `barrier()` orders those stores
    before the next reader (stream-ordered device-side barrier on the symmetric-memory signal pads)."""

    def __init__(self, rows: int, d: int, dtype: torch.dtype, device, group=None, multicast=None):
        import os
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.tensor = symm.empty((rows, d), dtype=dtype, device=device)
        self.handle = symm.rendezvous(self.tensor, self.group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        self.ptrs = [int(q) for q in self.handle.buffer_ptrs]
        self.row_bytes = d * self.tensor.element_size()
        # NVLS: one store to the multicast address is replicated by the NVSwitch into every rank's buffer, so a rank
        # sends each row ONCE instead of world-1 times (a multimem.st is an ordinary st.global on that address).
        self.multicast_ptr = 0
        if multicast is None:
            # a multicast store also comes back into the sender's own replica: 1x egress instead of (world-1)x, but
            # world/(world-1) of the ingress.  Measured: slower than per-peer stores at world 2 (2.71 vs 2.64 ms per step),
            # faster at 8 (1.13 vs 1.17).  ALLSET_MULTICAST=1 / 0 forces it on / off.
            env = os.environ.get('ALLSET_MULTICAST')
            multicast = (env == '1') if env in ('0', '1') else self.world >= 4
        if multicast:
            try:
                self.multicast_ptr = int(getattr(self.handle, 'multicast_ptr', 0) or 0)
            except Exception:  # noqa
                self.multicast_ptr = 0

    def peer_ptrs(self, first_row: int, unicast: bool = False):
        """Where the kernel must ALSO store row `first_row`...: the multicast address when the switch can replicate,
        else (or with `unicast`: a selective exchange addresses peers one by one) that row in every other rank's replica,
        in ascending rank order -- the bit order of `peer_need_mask`."""
        if self.multicast_ptr and not unicast:
            return [self.multicast_ptr + first_row * self.row_bytes]
        return [q + first_row * self.row_bytes for r, q in enumerate(self.ptrs) if r != self.rank]

    def barrier(self):
        self.handle.barrier(channel=0)


class ShardedIncidence(object):
    """One rank's share of a hypergraph: the hyperedge range it reduces in V->E and the vertex range it reduces in
    E->V, as slices of the two CSRs of a full `Incidence` (every rank holds the full index; features are what is
    big).  world == 1 degenerates to the single-GPU path with no collective.

        sh = ShardedIncidence(v2e, rank, world)
        sh.v2e_reduce(x_v, x_e)      # writes rows [e_lo, e_hi) of the replicated x_e
        sh.gather_e(x_e)             # all-gather of hyperedge rows
        sh.e2v_reduce(x_e, x_v_new)  # writes rows [v_lo, v_hi)
        sh.gather_v(x_v_new)         # all-gather of the updated vertex rows (north_star's one big collective)
    """

    def __init__(self, v2e, rank: int = 0, world: int = 1, group=None, row_weight: int = 1):
        from .graph import Csr
        self.rank, self.world, self.group = int(rank), int(world), group
        self.n_v, self.n_e, self.nnz = v2e.n_src, v2e.n_tgt, v2e.nnz
        t, s = v2e.by_tgt, v2e.by_src
        self.e_ranges = choose_ranges(t.rowptr, world, row_weight)
        self.v_ranges = choose_ranges(s.rowptr, world, row_weight)
        self.e_lo, self.e_hi = self.e_ranges[rank]
        self.v_lo, self.v_hi = self.v_ranges[rank]
        # (rows, longest segment) of EVERY rank's range, per direction: whether a direction takes the fused-exchange
        # kernel must be the same answer on all ranks (a rank that fell back to NCCL while its peers wait on the
        # symmetric-memory barrier would hang), so it is decided from these tables, never from the local slice alone.
        # Every rank holds the full rowptr, so no collective is needed to agree.
        self.e_shapes = range_shapes(t.rowptr, self.e_ranges)
        self.v_shapes = range_shapes(s.rowptr, self.v_ranges)
        self._need_masks = {}
        if world == 1:
            self.e_csr, self.v_csr = t, s
        else:
            rp, col, p0 = slice_csr(t.rowptr, t.col, self.e_lo, self.e_hi)
            self.e_csr = Csr(rp, col, t.perm[p0:p0 + col.numel()], self.e_hi - self.e_lo, self.n_v)
            rp, col, p0 = slice_csr(s.rowptr, s.col, self.v_lo, self.v_hi)
            self.v_csr = Csr(rp, col, s.perm[p0:p0 + col.numel()], self.v_hi - self.v_lo, self.n_e)

    def fused_ok(self, direction: str, x_src, out_full, heads: int = 0) -> bool:
        """Whether EVERY rank's range of `direction` ('e': V->E, 'v': E->V) is taken by the fused-exchange stream
        kernel for rows shaped like x_src (heads > 0: the PMA kernel).  Pure function of (graph, world, dtype, width):
        identical on all ranks."""
        from . import _lib
        if self.world == 1 or not isinstance(out_full, ReplicatedRows):
            return False
        shapes = self.e_shapes if direction == 'e' else self.v_shapes
        return all(_lib.fused_exchange_eligible(x_src.dtype, x_src.shape[1], rows, heads) for rows, _ in shapes)

    def need_mask(self, direction: str) -> torch.Tensor:
        """uint8 [rows of this rank's range]: bit j set = peer j (the j-th OTHER rank in ascending order) needs the row.
        direction 'v': vertex rows, needed by the ranks whose hyperedge range contains the vertex (what the next V->E
        gathers); 'e': hyperedge rows, needed by the ranks whose vertex range meets the hyperedge.  Built on first use
        from the CSR slice this rank already holds."""
        hit = self._need_masks.get(direction)
        if hit is None:
            if direction == 'v':
                csr, ranges = self.v_csr, self.e_ranges          # vertex -> hyperedges it belongs to
            else:
                csr, ranges = self.e_csr, self.v_ranges          # hyperedge -> member vertices
            hit = self._need_masks[direction] = peer_need_mask(csr.rowptr, csr.col, csr.n_tgt, ranges, self.rank)
        return hit

    # -- AllDeepSets ------------------------------------------------------------------------------------------
    def _reduce(self, direction, csr, x_src, out_full, lo, hi, mean, selective=False):
        """Reduce this rank's target range into rows [lo, hi) of the replicated buffer.  `out_full` is a plain tensor
        (exchange = a later all-gather) or a ReplicatedRows (exchange fused into the kernel's epilogue).  Returns True
        when the exchange has already been issued by the kernel -- the same value on every rank (`fused_ok`); a fused
        launch the library then rejects raises instead of silently diverging from the peers.  `selective`: send a row
        only to the peers that will gather it (`need_mask`; needs per-peer addresses, not the multicast one)."""
        from . import _lib
        if self.fused_ok(direction, x_src, out_full):
            mask = self.need_mask(direction) if selective else None
            _lib.segreduce_fwd_bcast(x_src, csr.rowptr, csr.col, csr.n_tgt, mean, out_full.tensor[lo:hi],
                                     out_full.peer_ptrs(lo, unicast=selective), peer_mask=mask)
            return True
        t = out_full.tensor if isinstance(out_full, ReplicatedRows) else out_full
        _lib.segreduce_fwd(x_src, csr.rowptr, csr.col, csr.n_tgt, mean, long_ids=csr.long_ids,
                           long_threshold=csr.long_threshold, out=t[lo:hi], max_segment_len=csr.max_len)
        return False

    def v2e_reduce(self, x_v, x_e_full, mean: bool = False, selective: bool = False):
        return self._reduce('e', self.e_csr, _plain(x_v), x_e_full, self.e_lo, self.e_hi, mean, selective)

    def e2v_reduce(self, x_e, x_v_full, mean: bool = False, selective: bool = False):
        return self._reduce('v', self.v_csr, _plain(x_e), x_v_full, self.v_lo, self.v_hi, mean, selective)

    # -- AllSetTransformer ------------------------------------------------------------------------------------
    def _pma(self, direction, csr, v, score, seed, heads, out_full, lo, hi, slope, selective=False):
        from . import _lib
        C = v.shape[1] // heads
        if self.fused_ok(direction, v, out_full, heads):
            mask = self.need_mask(direction) if selective else None
            _lib.pma_fwd_bcast(v, score, seed, heads, C, slope, csr.rowptr, csr.col, csr.n_tgt,
                               out_full.tensor[lo:hi], out_full.peer_ptrs(lo, unicast=selective), peer_mask=mask)
            return True
        t = out_full.tensor if isinstance(out_full, ReplicatedRows) else out_full
        _lib.pma_fwd(v, score, seed, heads, C, slope, csr.rowptr, csr.col, csr.n_tgt, want_stats=False,
                     long_ids=csr.long_ids, long_threshold=csr.long_threshold, out=t[lo:hi],
                     max_segment_len=csr.max_len)
        return False

    def v2e_pma(self, v_v, score_v, seed, heads, out_e_full, slope: float = 0.2, selective: bool = False):
        return self._pma('e', self.e_csr, _plain(v_v), score_v, seed, heads, out_e_full, self.e_lo, self.e_hi, slope,
                         selective)

    def e2v_pma(self, v_e, score_e, seed, heads, out_v_full, slope: float = 0.2, selective: bool = False):
        return self._pma('v', self.v_csr, _plain(v_e), score_e, seed, heads, out_v_full, self.v_lo, self.v_hi, slope,
                         selective)

    # -- exchanges --------------------------------------------------------------------------------------------
    def gather_e(self, x_e_full, fused: bool = False):
        """Complete the exchange of hyperedge rows: a barrier when the kernel already stored into the peers (fused),
        an NCCL all-gather otherwise."""
        self._gather(x_e_full, self.e_ranges, fused)

    def gather_v(self, x_v_full, fused: bool = False):
        self._gather(x_v_full, self.v_ranges, fused)

    def _gather(self, full, ranges, fused):
        if self.world == 1:
            return
        if fused and isinstance(full, ReplicatedRows):
            full.barrier()
        else:
            allgather_rows(full.tensor if isinstance(full, ReplicatedRows) else full, ranges, self.rank, self.group)

    def launches_per_pair(self) -> int:
        """Kernels of this library launched by one V->E + E->V pair on this rank."""
        return 2 + (self.e_csr.long_ids is not None) + (self.v_csr.long_ids is not None)

    def layer_pair_sum(self, x_v, x_e_full, x_v_new_full, mean: bool = False, replicate_v: bool = True):
        """V->E, exchange X_e, E->V and (replicate_v) exchange the updated X_v so that another layer can follow.
        With replicate_v=False the result stays vertex-sharded (rows [v_lo, v_hi) valid on this rank): enough when the
        next operator is row-parallel, e.g. SetGNN's classifier after the last layer."""
        self.gather_e(x_e_full, self.v2e_reduce(x_v, x_e_full, mean))
        if replicate_v:
            self.gather_v(x_v_new_full, self.e2v_reduce(x_e_full, x_v_new_full, mean))
        else:
            self.e2v_reduce(x_e_full, _plain(x_v_new_full), mean)
        return x_v_new_full

    def layer_pair_pma(self, v_v, score_v, score_e, seed, heads, out_e_full, out_v_full, replicate_v: bool = True):
        self.gather_e(out_e_full, self.v2e_pma(v_v, score_v, seed, heads, out_e_full))
        if replicate_v:
            self.gather_v(out_v_full, self.e2v_pma(out_e_full, score_e, seed, heads, out_v_full))
        else:
            self.e2v_pma(out_e_full, score_e, seed, heads, _plain(out_v_full))
        return out_v_full


def _plain(t):
    return t.tensor if isinstance(t, ReplicatedRows) else t
