"""`SetGNN` over a hypergraph partitioned across the GPUs of one node: one process per GPU, same module, same weights.

The reference is single-device; this is the scale-out BASELINE.json:north_star names ("the incidence
graph shards by hyperedge across GPUs with [an exchange] of updated vertex features per layer over NVLink").  Every
rank holds the model replica and the full (small) index; what is partitioned is the ROWS: rank r owns vertex rows
[v_lo, v_hi) and hyperedge rows [e_lo, e_hi) (`sharding.ShardedIncidence`).  All dense work (f_enc / f_dec MLPs, PMA's
projections and tail, classifier) is row-parallel on the owned rows; only the gather of a half layer needs rows of other
ranks, so each half layer is

    own source rows --publish--> replicated buffer --gather + reduce own target range--> own target rows

and its backward is the mirror image: the gradients of the own target rows are published, and each rank reduces, over
the TRANSPOSED slice it already holds for the other direction, the gradient of its own source rows -- no all-reduce /
reduce-scatter anywhere, because no segment is ever split across ranks (reference src/models.py:474-481 is the loop
being distributed; src/layers.py:633,145 the gathers).

`publish` = P2P stores of the owned rows into the peers' replicas over NVLink (symmetric memory) + one device-side
barrier.  Vertex rows go only to the ranks whose hyperedge range contains the vertex (`ShardedIncidence.need_mask`):
on a random graph that is ~half of the rows at 8 ranks, which halves the bytes of the one large (|V| x d) exchange.
Per layer and direction there are exactly two exchanges forward (|V| x d after V->E's f_enc, |E| x d after E->V's f_enc)
and two backward; the buffers alternate, so the barrier that completes one exchange also frees the buffer of the
previous one (no extra "done reading" barrier).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .sharding import ReplicatedRows, ShardedIncidence


class _LocalRows(object):
    """world == 1: the "replicated" table is just a tensor."""

    def __init__(self, tensor):
        self.tensor = tensor


class _Exchange(object):
    """Replicated row tables in symmetric memory, keyed by (side, role, width, dtype): side 'v' | 'e' is the row space,
    role 'fwd' | 'bwd' | a name keeps forward activations, gradients and side tables (scores, statistics) apart."""

    def __init__(self, sh: ShardedIncidence, device, group=None):
        self.sh, self.device, self.group = sh, device, group
        self.tables: Dict[Tuple, ReplicatedRows] = {}

    def table(self, side: str, role: str, width: int, dtype: torch.dtype) -> ReplicatedRows:
        key = (side, role, int(width), dtype)
        t = self.tables.get(key)
        if t is None:
            rows = self.sh.n_v if side == 'v' else self.sh.n_e
            if self.sh.world == 1:
                t = _LocalRows(torch.empty((rows, width), dtype=dtype, device=self.device))
            else:
                t = ReplicatedRows(rows, width, dtype, self.device, self.group, multicast=False)
            self.tables[key] = t
        return t

    def own(self, side: str) -> Tuple[int, int]:
        return (self.sh.v_lo, self.sh.v_hi) if side == 'v' else (self.sh.e_lo, self.sh.e_hi)

    def publish(self, side: str, tables_and_rows, selective: bool = True) -> None:
        """tables_and_rows: [(ReplicatedRows, rows [own range, width])]: copy the owned rows into the own replica, send
        them to the peers that gather them, then ONE barrier for the whole group of tables."""
        lo, hi = self.own(side)
        if self.sh.world == 1:
            for rep, rows in tables_and_rows:
                if rows.data_ptr() != rep.tensor[lo:hi].data_ptr():
                    rep.tensor[lo:hi].copy_(rows)
            return
        mask = self.sh.need_mask(side) if selective else None
        for rep, rows in tables_and_rows:
            view = rep.tensor[lo:hi]
            if rows.data_ptr() != view.data_ptr():
                view.copy_(rows)
            _lib.push_rows(view, rep.peer_ptrs(lo, unicast=True), mask)
        tables_and_rows[0][0].barrier()


class _ShardedSegReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_loc, w, d: 'ShardedDirection', mean: bool):
        ex, sh = d.ex, d.ex.sh
        x_loc = x_loc.contiguous()
        rep = ex.table(d.src_side, 'fwd', x_loc.shape[1], x_loc.dtype)
        ex.publish(d.src_side, [(rep, x_loc)], d.selective(d.src_side))
        csr = d.fwd_csr
        w_csr = None if w is None else w.detach().float().index_select(0, csr.perm64)
        out = _lib.segreduce_fwd(rep.tensor, csr.rowptr, csr.col, csr.n_tgt, mean, w=w_csr, long_ids=csr.long_ids,
                                 long_threshold=csr.long_threshold)
        ctx.d, ctx.mean = d, mean
        ctx.save_for_backward(None if w is None else w.detach())
        return out

    @staticmethod
    def backward(ctx, g_out):
        d, mean = ctx.d, ctx.mean
        (w,) = ctx.saved_tensors
        ex = d.ex
        g_out = g_out.contiguous()
        rep = ex.table(d.tgt_side, 'bwd', g_out.shape[1], g_out.dtype)
        ex.publish(d.tgt_side, [(rep, g_out)], d.selective(d.tgt_side))
        csr = d.bwd_csr                                   # segments = own source rows, columns = target rows (global)
        w_T = None if w is None else w.float().index_select(0, csr.perm64)
        g_x = _lib.segreduce_fwd(rep.tensor, csr.rowptr, csr.col, csr.n_tgt, False, w=w_T,
                                 src_scale=d.tgt_inv_count if mean else None, long_ids=csr.long_ids,
                                 long_threshold=csr.long_threshold)
        if ctx.needs_input_grad[1]:
            raise NotImplementedError('sharded SetGNN: gradients of per-incidence weights (LearnMask) are not partitioned')
        return g_x, None, None, None


class _ShardedPMA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v_loc, score_loc, seed, d: 'ShardedDirection', H: int, C: int, slope: float):
        ex = d.ex
        v_loc = v_loc.contiguous()
        score_loc = score_loc.float().contiguous()
        seed_f = seed.detach().float().reshape(-1).contiguous()
        rep_v = ex.table(d.src_side, 'fwd', H * C, v_loc.dtype)
        rep_s = ex.table(d.src_side, 'score', H, torch.float32)
        ex.publish(d.src_side, [(rep_v, v_loc), (rep_s, score_loc)], d.selective(d.src_side))
        csr = d.fwd_csr
        out, stats = _lib.pma_fwd(rep_v.tensor, rep_s.tensor, seed_f, H, C, slope, csr.rowptr, csr.col, csr.n_tgt,
                                  want_stats=True, long_ids=csr.long_ids, long_threshold=csr.long_threshold)
        ctx.d, ctx.H, ctx.C, ctx.slope = d, H, C, slope
        ctx.seed_shape, ctx.seed_dtype = seed.shape, seed.dtype
        ctx.save_for_backward(v_loc, score_loc, seed_f, out, stats)
        return out

    @staticmethod
    def backward(ctx, g_out):
        v_loc, score_loc, seed_f, out, stats = ctx.saved_tensors
        d, H, C, slope = ctx.d, ctx.H, ctx.C, ctx.slope
        ex = d.ex
        g_out = g_out.contiguous()
        grad_v = grad_score = grad_seed = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # the transposed kernel needs, for EVERY target row its sources touch: the gradient row, the softmax
            # statistics and D = <g, out - seed> -- three row tables over the target side, one barrier
            D = _lib.rowdot_heads(g_out, out, seed_f, H, C)
            rep_g = ex.table(d.tgt_side, 'bwd', H * C, g_out.dtype)
            rep_st = ex.table(d.tgt_side, 'stats', 2 * H, torch.float32)
            rep_D = ex.table(d.tgt_side, 'D', H, torch.float32)
            ex.publish(d.tgt_side, [(rep_g, g_out), (rep_st, stats.reshape(-1, 2 * H)), (rep_D, D)],
                       d.selective(d.tgt_side))
            csr = d.bwd_csr
            grad_v, grad_score = _lib.pma_bwd(rep_g.tensor, v_loc, score_loc, rep_st.tensor.view(-1, H, 2), rep_D.tensor,
                                              H, C, slope, csr.rowptr, csr.col, csr.n_tgt, long_ids=csr.long_ids,
                                              long_threshold=csr.long_threshold)
        if ctx.needs_input_grad[2]:
            # every rank holds a replica of the seed: its gradient is the sum over ALL target rows (all-reduced with the
            # other parameter gradients by the caller, see ShardedSetGNN.allreduce_gradients)
            grad_seed = g_out.float().sum(dim=0).reshape(ctx.seed_shape).to(ctx.seed_dtype)
        return grad_v, grad_score, grad_seed, None, None, None, None


class ShardedDirection(object):
    """One direction of message passing as a rank sees it: duck-types `Incidence` for the layers (`n_src`, `n_tgt`,
    `with_n_src`) and routes the two aggregation operators through the exchange."""

    def __init__(self, ex: _Exchange, direction: str, selective_v: bool = True):
        sh = ex.sh
        self.ex, self.direction = ex, direction
        if direction == 'v2e':
            self.src_side, self.tgt_side = 'v', 'e'
            self.fwd_csr, self.bwd_csr = sh.e_csr, sh.v_csr
            self.n_src, self.n_tgt = sh.v_hi - sh.v_lo, sh.e_hi - sh.e_lo
            full_tgt = sh.full.by_tgt
        else:
            self.src_side, self.tgt_side = 'e', 'v'
            self.fwd_csr, self.bwd_csr = sh.v_csr, sh.e_csr
            self.n_src, self.n_tgt = sh.e_hi - sh.e_lo, sh.v_hi - sh.v_lo
            full_tgt = sh.full.by_src
        self._full_tgt = full_tgt
        self._selective_v = selective_v
        self.nnz = self.fwd_csr.nnz

    def selective(self, side: str) -> bool:
        """Vertex rows are sent only to the ranks that gather them; hyperedge rows of the configured graphs are needed
        by (almost) every rank, so they go to all (no mask to read)."""
        return self._selective_v and side == 'v'

    @property
    def tgt_inv_count(self) -> torch.Tensor:
        return self._full_tgt.inv_count                   # 1 / max(len, 1) of EVERY target row (global ids)

    def with_n_src(self, n_src: int) -> 'ShardedDirection':
        if n_src != self.n_src:
            raise ValueError('sharded %s expects the %d source rows this rank owns, got %d' % (self.direction, self.n_src, n_src))
        return self

    def weights_all_one(self, norm) -> bool:
        if norm.is_floating_point():
            return False
        return bool((norm == 1).all().item()) if norm.numel() > 0 else True

    # the two operators (allset_b200.ops dispatches here)
    def sharded_segment_reduce(self, x, weight, reduce: str):
        return _ShardedSegReduce.apply(x, weight, self, reduce == 'mean')

    def sharded_pma_aggregate(self, v, score, seed, heads: int, slope: float):
        C = v.shape[1] // heads
        return _ShardedPMA.apply(v, score, seed, self, heads, C, float(slope))


class ShardedSetGNN(nn.Module):
    """Wraps a `SetGNN` replica: `forward(data)` takes the same `data` (full `x`, `edge_index`, `norm` on every rank, as
    the reference's `data.to(device)`) and returns the logits of the vertex rows THIS rank owns, [v_hi - v_lo, classes]
    (`gather_logits()` assembles the full matrix).  Gradients of the (replicated) parameters are partial sums over the
    owned rows: `allreduce_gradients()` completes them -- the only NCCL collective of a training step."""

    def __init__(self, model, group=None, selective: bool = True):
        super().__init__()
        import torch.distributed as dist
        self.model = model
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.selective = selective
        self._state = None

    def _directions(self, data):
        edge_index = data.edge_index
        st = self._state
        if st is not None and st[0] is edge_index and st[1] == edge_index._version:
            return st[2], st[3], st[4]
        n_nodes = int(getattr(data, 'num_nodes', None) or data.x.size(0))
        v2e, _ = self.model._graph(edge_index, n_nodes)
        sh = ShardedIncidence(v2e, self.rank, self.world, self.group)
        sh.full = v2e
        ex = _Exchange(sh, data.x.device, self.group)
        dv2e, de2v = ShardedDirection(ex, 'v2e', self.selective), ShardedDirection(ex, 'e2v', self.selective)
        self._state = (edge_index, edge_index._version, sh, dv2e, de2v)
        return sh, dv2e, de2v

    def forward(self, data):
        m = self.model
        if m.All_num_layers == 0 or m.GPR or m.LearnMask:
            raise NotImplementedError('ShardedSetGNN covers the plain layer stack (no GPR / LearnMask / classifier-only)')
        sh, dv2e, de2v = self._directions(data)
        # row-parallel from the first operator on; `data.x` is the full [N, F] matrix (as the reference's `data`) or, with
        # `data.num_nodes` set, just the rows this rank owns
        x = data.x if (getattr(data, 'num_nodes', None) and data.x.shape[0] == sh.v_hi - sh.v_lo) else data.x[sh.v_lo:sh.v_hi]
        norm = data.norm
        if norm is not None and not norm.is_floating_point():
            norm = None if dv2e.weights_all_one(norm) else norm
        x = F.dropout(x, p=0.2, training=m.training)                      # reference src/models.py:473
        for i, _ in enumerate(m.V2EConvs):
            x = m._half(m.V2EConvs[i], x, dv2e, norm)                       # :475-476 on the owned hyperedge rows
            x = m._half(m.E2VConvs[i], x, de2v, norm)                       # :478-479 on the owned vertex rows
        return m.classifier(x, out_dtype=torch.float32)                    # :482, row-parallel

    def gather_logits(self, local_logits: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        sh = self._state[2]
        if self.world == 1:
            return local_logits
        full = torch.empty((sh.n_v, local_logits.shape[1]), dtype=local_logits.dtype, device=local_logits.device)
        full[sh.v_lo:sh.v_hi] = local_logits
        from .sharding import allgather_rows
        allgather_rows(full, sh.v_ranges, self.rank, self.group)
        return full

    def allreduce_gradients(self) -> None:
        import torch.distributed as dist
        if self.world == 1:
            return
        grads = [p.grad for p in self.model.parameters() if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, group=self.group)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
