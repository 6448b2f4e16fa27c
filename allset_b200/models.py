"""`SetGNN` with the reference's module API (reference src/models.py:295-484): same constructor (`args`
namespace, optional `norm`), attribute / `state_dict` names (V2EConvs, E2VConvs, bnV2Es, bnE2Vs, classifier, MLP,
GPRweights, Importance), `reset_parameters()` and `forward(data) -> [N, num_classes]`, including its side effect
of zero-basing `data.edge_index[1]` in place.  Drop-in for reference src/train.py:28-42,437,462,478.

Below the API the V->E / E->V aggregations run as fused sm_100a kernels over an `Incidence` that is sorted once per
graph and cached on `data.edge_index` (the reference re-derives sizes with a device->host sync in every scatter).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Linear, Parameter

from . import ops
from .graph import Incidence, attach
from .layers import MLP, HalfNLHconv

__all__ = ['SetGNN']

_ATTR = '_allset_setgnn_graph'


class SetGNN(nn.Module):
    def __init__(self, args, norm=None, agg_dtype: Optional[torch.dtype] = None):
        """args: namespace with All_num_layers, dropout, aggregate, normalization, deepset_input_norm, GPR, LearnMask,
        num_features, MLP_hidden, MLP_num_layers, heads, PMA, Classifier_hidden, Classifier_num_layers, num_classes
        (reference src/train.py:221-289).  agg_dtype (extension): storage dtype of the rows the aggregation kernels
        gather, e.g. torch.bfloat16; None keeps the module dtype (fp32 = reference numerics)."""
        super().__init__()
        self.All_num_layers = args.All_num_layers
        self.dropout = args.dropout
        self.aggr = args.aggregate
        self.NormLayer = args.normalization
        self.InputNorm = args.deepset_input_norm
        self.GPR = args.GPR
        self.LearnMask = args.LearnMask
        self.V2EConvs = nn.ModuleList()
        self.E2VConvs = nn.ModuleList()
        self.bnV2Es = nn.ModuleList()      # created, reset and saved but never applied -- as in the reference
        self.bnE2Vs = nn.ModuleList()

        if self.LearnMask:
            self.Importance = Parameter(torch.ones(norm.size()))

        def classifier(in_channels):
            return MLP(in_channels=in_channels, hidden_channels=args.Classifier_hidden, out_channels=args.num_classes,
                       num_layers=args.Classifier_num_layers, dropout=self.dropout, Normalization=self.NormLayer,
                       InputNorm=False)

        def half(in_dim):
            return HalfNLHconv(in_dim=in_dim, hid_dim=args.MLP_hidden, out_dim=args.MLP_hidden,
                               num_layers=args.MLP_num_layers, dropout=self.dropout, Normalization=self.NormLayer,
                               InputNorm=self.InputNorm, heads=args.heads, attention=args.PMA)

        if self.All_num_layers == 0:
            self.classifier = classifier(args.num_features)
        else:
            for i in range(self.All_num_layers):
                self.V2EConvs.append(half(args.num_features if i == 0 else args.MLP_hidden))
                self.bnV2Es.append(nn.BatchNorm1d(args.MLP_hidden))
                self.E2VConvs.append(half(args.MLP_hidden))
                self.bnE2Vs.append(nn.BatchNorm1d(args.MLP_hidden))
            if self.GPR:
                self.MLP = MLP(in_channels=args.num_features, hidden_channels=args.MLP_hidden,
                               out_channels=args.MLP_hidden, num_layers=args.MLP_num_layers, dropout=self.dropout,
                               Normalization=self.NormLayer, InputNorm=False)
                self.GPRweights = Linear(self.All_num_layers + 1, 1, bias=False)
            self.classifier = classifier(args.MLP_hidden)
        self.set_agg_dtype(agg_dtype)

    def set_agg_dtype(self, dtype: Optional[torch.dtype]):
        self.agg_dtype = dtype
        for conv in list(self.V2EConvs) + list(self.E2VConvs):
            conv.set_agg_dtype(dtype)
        tc = dtype if dtype == torch.bfloat16 else None          # bf16 mode: the classifier's Linears take bf16 operands too
        self.classifier.tc_dtype = tc
        if self.GPR and self.All_num_layers > 0:
            self.MLP.tc_dtype = tc

    def reset_parameters(self):
        for group in (self.V2EConvs, self.E2VConvs, self.bnV2Es, self.bnE2Vs):
            for layer in group:
                layer.reset_parameters()
        self.classifier.reset_parameters()
        if self.GPR:
            self.MLP.reset_parameters()
            self.GPRweights.reset_parameters()
        if self.LearnMask:
            nn.init.ones_(self.Importance)

    # ------------------------------------------------------------------------------------------------------
    def _graph(self, edge_index: torch.Tensor, n_nodes: int):
        """(V->E incidence, E->V incidence) for data.edge_index, built on first sight and cached on the tensor.

        Reproduces reference src/models.py:453-454: hyperedge ids are zero-based IN PLACE (cidx = row1.min();
        row1 -= cidx).  The reference repeats the subtraction every forward (a no-op after the first); here the
        shift happens when the graph is (re)built, and later forwards reuse the cache with no device->host sync."""
        if not edge_index.is_cuda:
            raise RuntimeError('allset_b200.SetGNN runs on CUDA only (no CPU fallback): data.edge_index is on %s'
                               % edge_index.device)
        tagged = getattr(edge_index, _ATTR, None)
        if tagged is not None and tagged[0] == edge_index._version and tagged[1] == n_nodes:
            return tagged[2], tagged[3]
        if edge_index.numel() > 0:
            cidx = edge_index[1].min()
            edge_index[1] -= cidx
        v2e = Incidence.from_coo(edge_index[0], edge_index[1], n_src=n_nodes)          # rows out: max(he)+1
        n_v_out = int(edge_index[0].max()) + 1 if edge_index.numel() > 0 else 0        # rows out of E->V: max(node)+1
        e2v = v2e.reversed(n_tgt=n_v_out)
        setattr(edge_index, _ATTR, (edge_index._version, n_nodes, v2e, e2v))
        return v2e, e2v

    def _half(self, conv, x, inc, norm):
        """`F.dropout(F.relu(conv(x, ...)))` of reference src/models.py:475-476,478-479 with the ReLU and the dropout
        folded into the layer's last fused pass.  A non-attention half layer already ends in relu(f_dec(.))
        (src/layers.py:634): the outer ReLU is the identity (value and gradient).  A PMA half layer ends in a LayerNorm:
        the ReLU is applied by its tail kernel (a forward hook on the layer then sees post-ReLU, post-dropout rows)."""
        return conv(x, inc, norm, self.aggr, relu_out=True, out_dropout=self.dropout)

    def forward(self, data):
        """data.x [N, F]; data.edge_index [2, nnz] int64 (row 0 node id, row 1 hyperedge id, any base);
        data.norm [nnz] per-incidence weights (int64 ones by default).  Returns node logits [N, num_classes]."""
        x, edge_index, norm = data.x, data.edge_index, data.norm
        if self.All_num_layers == 0:
            # only the classifier exists (src/models.py:339-346); the reference still zero-bases the hyperedge ids
            if edge_index.numel() > 0:
                edge_index[1] -= edge_index[1].min()
            return self.classifier(F.dropout(x, p=0.2, training=self.training))
        if self.LearnMask:
            norm = self.Importance * norm
        v2e, e2v = self._graph(edge_index, x.size(0))
        if self.GPR:
            xs = [F.relu(self.MLP(x))]
            for i, _ in enumerate(self.V2EConvs):
                x = self.V2EConvs[i](x, v2e, norm, self.aggr, relu_out=True).float()     # = F.relu(conv(.)), :460
                x = F.dropout(x, p=self.dropout, training=self.training)
                x = self.E2VConvs[i](x, e2v, norm, self.aggr, relu_out=True).float()     # :464
                xs.append(x)
                x = F.dropout(x, p=self.dropout, training=self.training)
            x = torch.stack(xs, dim=-1)
            x = self.GPRweights(x).squeeze()
            x = self.classifier(x, out_dtype=torch.float32)
        else:
            if self.training and self.agg_dtype == torch.bfloat16 and torch.is_grad_enabled() and ops.rowop_ok(x) \
                    and x.shape[0] >= ops.FUSED_DENSE_MIN_ROWS:
                x = ops.rowop(x, drop_p=0.2, out_dtype=torch.bfloat16)   # input dropout + cast to the bf16 mode's rows, one pass
            else:
                x = F.dropout(x, p=0.2, training=self.training)     # input dropout, hard-coded in the reference
            for i, _ in enumerate(self.V2EConvs):
                x = self._half(self.V2EConvs[i], x, v2e, norm)                    # = dropout(relu(conv(.))), :475-476
                x = self._half(self.E2VConvs[i], x, e2v, norm)                    # :478-479
            x = self.classifier(x, out_dtype=torch.float32)
        return x
