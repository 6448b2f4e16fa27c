"""B200-native `UniGCNIIConv` / `UniGCNII` (reference src/models.py:909-995): the comparison model of the reference
that shares the hot path's shape -- per layer a V->E MEAN followed by an E->V SUM over the same incidence list, with
per-hyperedge / per-vertex degree scales -- so it runs on the same two segmented-reduce kernels as AllDeepSets
(SURVEY.md 8f-3).  Same constructors, attribute names, `state_dict` keys and forward as the reference, so
`train.py --method UniGCNII` (reference src/train.py:92-103,390-418) drives it unchanged through the drop-in `models`.

Reference per layer (models.py:919-942):
    Xe = scatter(X[vertex], edges, reduce='mean') * degE        # [E, C]
    Xv = scatter(Xe[edges], vertex, reduce='sum', dim_size=N) * degV
    X  = normalize_l2(Xv) if use_norm;  Xi = (1-alpha) X + alpha X0;  X = (1-beta) Xi + beta W(Xi)
Here the two scatter chains (2 gathers + 2 atomic scatters + count pass, five [nnz, C] round trips) are two launches of
`segment_reduce` over an `Incidence` built once from (V, E); the rest stays elementwise ATen / cuBLAS as in the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .graph import Incidence

__all__ = ['UniGCNIIConv', 'UniGCNII', 'normalize_l2']


def normalize_l2(X):
    """Row-normalise (reference src/models.py:590-596): rows of zero norm stay zero."""
    rownorm = X.detach().norm(dim=1, keepdim=True)
    scale = rownorm.pow(-1)
    scale[torch.isinf(scale)] = 0.
    return X * scale


def _incidence(vertex: torch.Tensor, edges: torch.Tensor, n_nodes: int) -> Incidence:
    """(vertex, edges) -> cached Incidence (V->E by target hyperedge; .reversed() for E->V), keyed on the tensors."""
    if not vertex.is_cuda:
        raise RuntimeError('allset_b200.UniGCNII runs on CUDA only (no CPU fallback): V is on %s' % vertex.device)
    tag = getattr(vertex, '_allset_uni_graph', None)
    if tag is not None and tag[0] is edges and tag[1] == (vertex._version, edges._version, n_nodes):
        return tag[2], tag[3]
    v2e = Incidence.from_coo(vertex, edges, n_src=n_nodes)                  # rows out: edges.max()+1, as scatter()
    e2v = v2e.reversed(n_tgt=n_nodes)                                       # dim_size=N
    vertex._allset_uni_graph = (edges, (vertex._version, edges._version, n_nodes), v2e, e2v)
    return v2e, e2v


class UniGCNIIConv(nn.Module):
    def __init__(self, args, in_features, out_features):
        super().__init__()
        self.W = nn.Linear(in_features, out_features, bias=False)
        self.args = args

    def reset_parameters(self):
        self.W.reset_parameters()

    def forward(self, X, vertex, edges, alpha, beta, X0):
        N = X.shape[0]
        degE = self.args.UniGNN_degE
        degV = self.args.UniGNN_degV
        v2e, e2v = _incidence(vertex, edges, N)
        Xe = ops.segment_reduce(X, v2e, None, 'mean')          # scatter(X[vertex], edges, reduce='mean')
        Xe = Xe * degE
        Xv = ops.segment_reduce(Xe, e2v, None, 'sum')           # scatter(Xe[edges], vertex, reduce='sum', dim_size=N)
        Xv = Xv * degV
        X = Xv
        if self.args.UniGNN_use_norm:
            X = normalize_l2(X)
        Xi = (1 - alpha) * X + alpha * X0
        X = (1 - beta) * Xi + beta * self.W(Xi)
        return X


class UniGCNII(nn.Module):
    def __init__(self, args, nfeat, nhid, nclass, nlayer, nhead, V, E):
        """Same signature as the reference (models.py:949-963): V / E are the row / column indices of the incidence
        matrix H [|V| x |E|]."""
        super().__init__()
        self.V = V
        self.E = E
        nhid = nhid * nhead
        self.act = nn.ReLU()
        self.input_drop = nn.Dropout(0.6)
        self.dropout = nn.Dropout(0.2)
        self.convs = torch.nn.ModuleList()
        self.convs.append(torch.nn.Linear(nfeat, nhid))
        for _ in range(nlayer):
            self.convs.append(UniGCNIIConv(args, nhid, nhid))
        self.convs.append(torch.nn.Linear(nhid, nclass))
        self.reg_params = list(self.convs[1:-1].parameters())
        self.non_reg_params = list(self.convs[0:1].parameters()) + list(self.convs[-1:].parameters())

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()

    def forward(self, data):
        x = data.x
        V, E = self.V, self.E
        lamda, alpha = 0.5, 0.1
        x = self.dropout(x)
        x = F.relu(self.convs[0](x))
        x0 = x
        for i, con in enumerate(self.convs[1:-1]):
            x = self.dropout(x)
            beta = math.log(lamda / (i + 1) + 1)
            x = F.relu(con(x, V, E, alpha, beta, x0))
        x = self.dropout(x)
        x = self.convs[-1](x)
        return x
