"""The reference's comparison convolutions that share the hot path's shape -- a gather over the incidence list followed by
a segmented reduce per hyperedge, then the same per vertex -- on the SAME two kernels as AllDeepSets / AllSetTransformer
(SURVEY.md 8f-3).  Constructors, attribute names, `state_dict` keys and forward signatures are the reference's, so its
`train.py` drives them unchanged through the drop-in `layers` / `models` modules:

    HypergraphConv, HCHA      reference src/layers.py:318-494, src/models.py:252-292   (HGNN / HCHA: mean, mean)
    HNHNConv, HNHN            reference src/layers.py:233-315, src/models.py:207-249   (two weighted sums)
    UniSAGEConv, UniGINConv, UniGCNConv, UniGCNConv2, UniGATConv, UniGNN
                              reference src/models.py:601-907                           (UniGAT = segment softmax)

What the reference does per direction -- `X[vertex]` (a materialised [nnz, C] gather), `torch_scatter.scatter` (atomics),
for UniGAT also PyG's 6-kernel segment softmax -- is ONE launch of `segment_reduce` / `pma_aggregate` over an `Incidence`
sorted once per graph.  Per-row degree scales and the Linears stay elementwise ATen / cuBLAS, as in the reference.
Reductions the kernels do not have ('max' / 'min' as `first_aggregate`) are composed from ATen scatter ops on the same
device.  No CPU path.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn import Linear, Parameter

from . import ops
from .graph import Incidence
from .layers import glorot, zeros
from .uni import normalize_l2

__all__ = ['HypergraphConv', 'HCHA', 'HNHNConv', 'HNHN', 'UniSAGEConv', 'UniGINConv', 'UniGCNConv', 'UniGCNConv2',
           'UniGATConv', 'UniGNN']


def _pair(src_ids: Tensor, tgt_ids: Tensor, n_src: int, n_tgt: Optional[int] = None, holder: Optional[Tensor] = None):
    """(source -> target incidence, its reverse) for two id vectors, sorted once and cached on `holder` (the tensor
    object that persists across forwards: `data.edge_index` itself, or the model's `V`), keyed on its version."""
    if not src_ids.is_cuda:
        raise RuntimeError('allset_b200 baselines run on CUDA only (no CPU fallback): the incidence list is on %s'
                           % src_ids.device)
    holder = src_ids if holder is None else holder
    key = (holder._version, tgt_ids.data_ptr(), tgt_ids._version, n_src, n_tgt)
    tag = getattr(holder, '_allset_pair', None)
    if tag is not None and tag[0] == key:
        return tag[1], tag[2]
    fwd = Incidence.from_coo(src_ids, tgt_ids, n_src=n_src, n_tgt=n_tgt)       # rows out: tgt.max()+1 unless given
    rev = fwd.reversed(n_tgt=n_src)
    try:
        holder._allset_pair = (key, fwd, rev)
    except Exception:  # noqa
        pass
    return fwd, rev


def _reduce(x: Tensor, inc: Incidence, reduce: str, src_ids: Tensor, tgt_ids: Tensor, n_tgt: int) -> Tensor:
    """torch_scatter.scatter(x[src_ids], tgt_ids, dim=0, reduce=...) -- sum / mean on the segmented-reduce kernel, the
    rest composed from ATen scatter ops (same device)."""
    if reduce in ('sum', 'add', 'mean'):
        return ops.segment_reduce(x, inc, None, reduce)
    if reduce in ('max', 'min'):
        out = x.new_zeros((n_tgt,) + tuple(x.shape[1:]))
        idx = tgt_ids.view((-1,) + (1,) * (x.dim() - 1)).expand(-1, *x.shape[1:])
        return out.scatter_reduce(0, idx, x.index_select(0, src_ids), reduce='a' + reduce, include_self=False)
    raise ValueError('unknown reduce %r' % (reduce,))


# ----------------------------------------------------------------------------------------------------------------------
# HGNN / HCHA
# ----------------------------------------------------------------------------------------------------------------------
class HypergraphConv(nn.Module):
    """X' = D^-1 H W B^-1 H^T X Theta (reference src/layers.py:318-494).  Without attention (all HCHA / HGNN uses) this is
    a V->E MEAN followed by an E->V MEAN (`symdegnorm`: D^-1/2 on both sides of an E->V SUM instead).
    `hyperedge_index` row 0 = node ids, row 1 = hyperedge ids (zero-based, as train.py:383 prepares them)."""

    def __init__(self, in_channels, out_channels, symdegnorm=False, use_attention=False, heads=1, concat=True,
                 negative_slope=0.2, dropout=0, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_attention, self.symdegnorm = use_attention, symdegnorm
        if use_attention:
            self.heads, self.concat, self.negative_slope, self.dropout = heads, concat, negative_slope, dropout
            self.weight = Parameter(torch.Tensor(in_channels, heads * out_channels))
            self.att = Parameter(torch.Tensor(1, heads, 2 * out_channels))
        else:
            self.heads, self.concat = 1, True
            self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.weight)
        if self.use_attention:
            glorot(self.att)
        zeros(self.bias)

    def forward(self, x: Tensor, hyperedge_index: Tensor, hyperedge_weight: Optional[Tensor] = None) -> Tensor:
        n = x.size(0)
        node, he = hyperedge_index[0], hyperedge_index[1]
        m = int(he.max()) + 1 if hyperedge_index.numel() > 0 else 0
        x = torch.matmul(x, self.weight)
        if self.use_attention:
            return self._forward_attention(x, node, he, n, m, hyperedge_weight)
        v2e, e2v = _pair(node, he, n, m, holder=hyperedge_index)
        cnt_e = (v2e.by_tgt.rowptr[1:m + 1] - v2e.by_tgt.rowptr[:m]).to(x.dtype)
        B = torch.where(cnt_e > 0, 1.0 / cnt_e, torch.zeros_like(cnt_e))                 # 1 / |e|, inf -> 0
        if hyperedge_weight is None:
            deg = (e2v.by_tgt.rowptr[1:n + 1] - e2v.by_tgt.rowptr[:n]).to(x.dtype)       # hyperedge_weight == 1
        else:
            deg = ops.segment_reduce(hyperedge_weight.view(-1, 1).to(x.dtype), e2v, None, 'sum').view(-1)
        D = torch.where(deg != 0, deg.pow(-0.5 if self.symdegnorm else -1.0), torch.zeros_like(deg))
        if self.symdegnorm:
            x = D.unsqueeze(-1) * x
        out = ops.segment_reduce(x, v2e, None, 'sum') * B.unsqueeze(-1)                  # norm_i = B[e]
        out = ops.segment_reduce(out, e2v, None, 'sum') * D.unsqueeze(-1)                # norm_i = D[v]
        return out if self.bias is None else out + self.bias

    def _forward_attention(self, x, node, he, n, m, hyperedge_weight):
        """Attention variant (unused by the reference's own models): per-incidence coefficients, composed from ATen
        gather / index_add ops around the same propagate algebra (src/layers.py:425-434,480-489)."""
        H, Fo = self.heads, self.out_channels
        x = x.view(-1, H, Fo)
        x_i, x_j = x[node], x[he]                                     # the reference indexes NODE rows by hyperedge id
        alpha = F.leaky_relu((torch.cat([x_i, x_j], dim=-1) * self.att).sum(dim=-1), self.negative_slope)
        amax = alpha.new_full((n, H), float('-inf')).scatter_reduce(0, node.view(-1, 1).expand(-1, H), alpha, 'amax')
        alpha = (alpha - amax[node]).exp()
        alpha = alpha / (alpha.new_zeros((n, H)).index_add_(0, node, alpha)[node] + 1e-16)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        w = x.new_ones(m) if hyperedge_weight is None else hyperedge_weight
        deg = x.new_zeros(n).index_add_(0, node, w[he])
        D = torch.where(deg != 0, deg.pow(-0.5 if self.symdegnorm else -1.0), torch.zeros_like(deg))
        cnt = x.new_zeros(m).index_add_(0, he, x.new_ones(he.numel()))
        B = torch.where(cnt != 0, 1.0 / cnt, torch.zeros_like(cnt))
        if self.symdegnorm:
            x = D.unsqueeze(-1) * x              # as in the reference (:464): does not broadcast over [n, H, F] -- it raises
        msg = alpha.unsqueeze(-1) * (B[he].view(-1, 1, 1) * x[node])
        out = x.new_zeros((m, H, Fo)).index_add_(0, he, msg)
        msg = alpha.unsqueeze(-1) * (D[node].view(-1, 1, 1) * out[he])
        out = x.new_zeros((n, H, Fo)).index_add_(0, node, msg)
        out = out.view(-1, H * Fo) if self.concat else out.mean(dim=1)
        return out if self.bias is None else out + self.bias

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class HCHA(nn.Module):
    """Stack of HypergraphConv layers with ELU + dropout between them (reference src/models.py:252-292)."""

    def __init__(self, args):
        super().__init__()
        self.num_layers = args.All_num_layers
        self.dropout = args.dropout
        self.symdegnorm = args.HCHA_symdegnorm
        widths = [args.num_features] + [args.MLP_hidden] * (max(self.num_layers, 2) - 1) + [args.num_classes]
        self.convs = nn.ModuleList(HypergraphConv(a, b, self.symdegnorm) for a, b in zip(widths[:-1], widths[1:]))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, data):
        x, edge_index = data.x, data.edge_index
        last = len(self.convs) - 1
        for i, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if i < last:
                x = F.dropout(F.elu(x), p=self.dropout, training=self.training)
        return x


# ----------------------------------------------------------------------------------------------------------------------
# HNHN
# ----------------------------------------------------------------------------------------------------------------------
class HNHNConv(nn.Module):
    """Hyperedge neurons (reference src/layers.py:233-315): Linear, D_v^beta scale, V->E sum scaled by D_e_beta_inv,
    [ReLU], Linear, D_e^alpha scale, E->V sum scaled by D_v_alpha_inv.  `data` carries edge_index (row 0 nodes, row 1
    zero-based hyperedges) and the four degree vectors of `generate_norm_HNHN` (reference src/preprocessing.py)."""

    def __init__(self, in_channels, hidden_channels, out_channels, heads=1, nonlinear_inbetween=True, concat=True,
                 bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
        self.nonlinear_inbetween = nonlinear_inbetween
        self.heads, self.concat = heads, True
        self.weight_v2e = Linear(in_channels, hidden_channels)
        self.weight_e2v = Linear(hidden_channels, out_channels)
        self.reset_parameters()

    def reset_parameters(self):
        self.weight_v2e.reset_parameters()
        self.weight_e2v.reset_parameters()

    def forward(self, x, data):
        edge_index = data.edge_index
        n = x.size(0)
        m = int(edge_index[1].max()) + 1 if edge_index.numel() > 0 else 0
        v2e, e2v = _pair(edge_index[0], edge_index[1], n, m, holder=edge_index)
        x = data.D_v_beta.unsqueeze(-1) * self.weight_v2e(x)
        out = ops.segment_reduce(x, v2e, None, 'sum') * data.D_e_beta_inv.view(-1, 1)     # message: norm_i * x_j
        if self.nonlinear_inbetween:
            out = F.relu(out)
        out = data.D_e_alpha.unsqueeze(-1) * self.weight_e2v(torch.squeeze(out, dim=1))
        return ops.segment_reduce(out, e2v, None, 'sum') * data.D_v_alpha_inv.view(-1, 1)

    def __repr__(self):
        return '{}({}, {}, {})'.format(self.__class__.__name__, self.in_channels, self.hidden_channels, self.out_channels)


class HNHN(nn.Module):
    """reference src/models.py:207-249."""

    def __init__(self, args):
        super().__init__()
        self.num_layers = args.All_num_layers
        self.dropout = args.dropout
        nl = args.HNHN_nonlinear_inbetween
        if self.num_layers == 1:
            dims = [(args.num_features, args.MLP_hidden, args.num_classes)]
        else:
            dims = [(args.num_features, args.MLP_hidden, args.MLP_hidden)]
            dims += [(args.MLP_hidden, args.MLP_hidden, args.MLP_hidden)] * (self.num_layers - 2)
            dims += [(args.MLP_hidden, args.MLP_hidden, args.num_classes)]
        self.convs = nn.ModuleList(HNHNConv(a, b, c, nonlinear_inbetween=nl) for a, b, c in dims)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, data):
        x = data.x
        for conv in self.convs[:-1]:
            x = F.dropout(F.relu(conv(x, data)), p=self.dropout, training=self.training)
        return self.convs[-1](x, data)


# ----------------------------------------------------------------------------------------------------------------------
# UniGNN family
# ----------------------------------------------------------------------------------------------------------------------
class _UniConv(nn.Module):
    """Shared constructor of the UniGNN convolutions (reference src/models.py:601-612 and its copies)."""

    def __init__(self, args, in_channels, out_channels, heads=8, dropout=0., negative_slope=0.2, bias=False):
        super().__init__()
        self.W = nn.Linear(in_channels, heads * out_channels, bias=bias)
        self.heads, self.in_channels, self.out_channels = heads, in_channels, out_channels
        self.negative_slope, self.dropout, self.args = negative_slope, dropout, args

    def __repr__(self):
        return '{}({}, {}, heads={})'.format(self.__class__.__name__, self.in_channels, self.out_channels, self.heads)

    def _two_hops(self, X, vertex, edges, second: str, scale_e=None, scale_v=None):
        """Xe = scatter(X[vertex], edges, first_aggregate) [* degE];  Xv = scatter(Xe[edges], vertex, second, N) [* degV]"""
        N = X.shape[0]
        v2e, e2v = _pair(vertex, edges, N)
        Xe = _reduce(X, v2e, self.args.first_aggregate, vertex, edges, v2e.n_tgt)
        if scale_e is not None:
            Xe = Xe * scale_e
        Xv = _reduce(Xe, e2v, second, edges, vertex, N)
        return Xv if scale_v is None else Xv * scale_v


class UniSAGEConv(_UniConv):
    def forward(self, X, vertex, edges):                       # reference src/models.py:619-640
        X = self.W(X)
        X = X + self._two_hops(X, vertex, edges, self.args.second_aggregate)
        return normalize_l2(X) if self.args.use_norm else X


class UniGINConv(_UniConv):
    def __init__(self, args, in_channels, out_channels, heads=8, dropout=0., negative_slope=0.2):
        super().__init__(args, in_channels, out_channels, heads, dropout, negative_slope)
        self.eps = nn.Parameter(torch.Tensor([0.]))

    def forward(self, X, vertex, edges):                       # reference src/models.py:665-687
        X = self.W(X)
        X = (1 + self.eps) * X + self._two_hops(X, vertex, edges, 'sum')
        return normalize_l2(X) if self.args.use_norm else X


class UniGCNConv(_UniConv):
    def forward(self, X, vertex, edges):                       # reference src/models.py:711-736
        X = self._two_hops(self.W(X), vertex, edges, 'sum', self.args.degE, self.args.degV)
        return normalize_l2(X) if self.args.use_norm else X


class UniGCNConv2(_UniConv):
    def __init__(self, args, in_channels, out_channels, heads=8, dropout=0., negative_slope=0.2):
        super().__init__(args, in_channels, out_channels, heads, dropout, negative_slope, bias=True)

    def forward(self, X, vertex, edges):                       # reference src/models.py:759-787: aggregate, then W
        X = self._two_hops(X, vertex, edges, 'sum', self.args.degE, self.args.degV)
        if self.args.use_norm:
            X = normalize_l2(X)
        return self.W(X)


class UniGATConv(_UniConv):
    """E->V with attention (reference src/models.py:791-846): score of hyperedge e and head h = <Xe[e,h,:], att_e[h,:]>,
    softmax of leaky_relu(score) over the hyperedges of each vertex, weighted sum -- the PMA kernel with the scores given
    per source row and a zero seed.  `att_v` is a parameter of the reference that its forward never reads."""

    def __init__(self, args, in_channels, out_channels, heads=8, dropout=0., negative_slope=0.2, skip_sum=False):
        super().__init__(args, in_channels, out_channels, heads, dropout, negative_slope)
        self.att_v = nn.Parameter(torch.Tensor(1, heads, out_channels))
        self.att_e = nn.Parameter(torch.Tensor(1, heads, out_channels))
        self.attn_drop = nn.Dropout(dropout)
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        self.skip_sum = skip_sum
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.att_v)
        glorot(self.att_e)

    def forward(self, X, vertex, edges):
        H, C, N = self.heads, self.out_channels, X.shape[0]
        X0 = self.W(X)
        v2e, e2v = _pair(vertex, edges, N)
        Xe = _reduce(X0, v2e, self.args.first_aggregate, vertex, edges, v2e.n_tgt)          # [E, H*C]
        alpha_e = (Xe.view(-1, H, C) * self.att_e).sum(-1)                                   # [E, H]
        if self.training and self.attn_drop.p > 0:
            # dropout on the attention coefficients needs them materialised per incidence: ATen composition
            a = self.leaky_relu(alpha_e[edges])
            amax = a.new_full((N, H), float('-inf')).scatter_reduce(0, vertex.view(-1, 1).expand(-1, H), a, 'amax')
            a = (a - amax[vertex]).exp()
            a = self.attn_drop(a / (a.new_zeros((N, H)).index_add_(0, vertex, a)[vertex] + 1e-16))
            Xv = Xe.new_zeros((N, H, C)).index_add_(0, vertex, Xe.view(-1, H, C)[edges] * a.unsqueeze(-1)).view(N, H * C)
        else:
            seed = Xe.new_zeros((1, H, C))
            Xv, _ = ops.pma_aggregate(Xe, alpha_e, seed, e2v, H, self.negative_slope)
        X = normalize_l2(Xv) if self.args.use_norm else Xv
        return X + X0 if self.skip_sum else X


_UNI_CONVS = {'UniGAT': UniGATConv, 'UniGCN': UniGCNConv, 'UniGCN2': UniGCNConv2, 'UniGIN': UniGINConv,
              'UniSAGE': UniSAGEConv}


class UniGNN(nn.Module):
    """reference src/models.py:861-907: `nlayer - 1` hidden convolutions + an output convolution with one head; V / E are
    the row / column indices of the incidence matrix."""

    def __init__(self, args, nfeat, nhid, nclass, nlayer, nhead, V, E):
        super().__init__()
        Conv = _UNI_CONVS[args.model_name]
        self.conv_out = Conv(args, nhid * nhead, nclass, heads=1, dropout=args.attn_drop)
        ins = [nfeat] + [nhid * nhead] * (nlayer - 2)
        self.convs = nn.ModuleList(Conv(args, i, nhid, heads=nhead, dropout=args.attn_drop) for i in ins)
        self.V, self.E = V, E
        self.act = {'relu': nn.ReLU(), 'prelu': nn.PReLU()}[args.activation]
        self.input_drop = nn.Dropout(args.input_drop)
        self.dropout = nn.Dropout(args.dropout)

    def forward(self, X):
        X = self.input_drop(X)
        for conv in self.convs:
            X = self.dropout(self.act(conv(X, self.V, self.E)))
        return F.log_softmax(self.conv_out(X, self.V, self.E), dim=1)
