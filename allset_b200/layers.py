"""B200-native `MLP`, `PMA` and `HalfNLHconv` with the reference's constructor signatures, attribute names,
initialisation and `state_dict` keys (reference src/layers.py: PMA :42-199, MLP :496-579, HalfNLHconv :582-656),
so reference checkpoints load and reference `train.py` drives them unchanged.

What differs is below the module API: instead of PyG `MessagePassing.propagate` + torch_scatter (gather ->
materialised [nnz, d] messages -> atomic scatter, ~20 launches for PMA) each half layer issues ONE fused CUDA
kernel over a cached CSR (allset_b200/ops.py).  The dense parts (Linear / LayerNorm) stay on cuBLAS / ATen.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor
from torch.nn import Linear, Parameter

from . import _lib, ops
from .graph import Incidence, incidence_of

__all__ = ['MLP', 'PMA', 'HalfNLHconv', 'glorot', 'zeros']


def glorot(tensor):
    """Uniform(-a, a), a = sqrt(6 / (fan_in + fan_out)) over the last two dims (reference src/layers.py:31-34)."""
    if tensor is not None:
        bound = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-bound, bound)


def zeros(tensor):
    if tensor is not None:
        tensor.data.fill_(0)


class MLP(nn.Module):
    """norm0 -> [Linear -> ReLU -> norm -> dropout] x (num_layers-1) -> Linear   (reference src/layers.py:496-579).

    `normalizations[0]` is a real norm only when `InputNorm`; Normalization 'None' makes every norm Identity."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers,
                 dropout=.5, Normalization='bn', InputNorm=False):
        super().__init__()
        assert Normalization in ['bn', 'ln', 'None']
        self.InputNorm = InputNorm
        self.dropout = dropout
        make = {'bn': nn.BatchNorm1d, 'ln': nn.LayerNorm, 'None': lambda width: nn.Identity()}[Normalization]
        # the reference's ctor builds in->hid, (num_layers-2) x hid->hid, hid->out for every num_layers != 1
        # (so num_layers <= 0 still yields two Linears)
        depth = 1 if num_layers == 1 else max(num_layers, 2)
        widths = [in_channels] + [hidden_channels] * (depth - 1) + [out_channels]
        self.lins = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))
        norms = [make(in_channels) if (InputNorm and Normalization != 'None') else nn.Identity()]
        norms += [make(hidden_channels) for _ in range(depth - 1)]
        self.normalizations = nn.ModuleList(norms)
        # extension: torch.bfloat16 lets eval-mode forwards of square two-layer MLPs run as ONE tcgen05 kernel with
        # bf16 operands (allset_mlp2_fwd); None (default) keeps fp32 numerics (cuBLAS SGEMM + fused glue)
        self.tc_dtype: Optional[torch.dtype] = None

    def reset_parameters(self):
        for lin in self.lins:
            lin.reset_parameters()
        for norm in self.normalizations:
            if not isinstance(norm, nn.Identity):
                norm.reset_parameters()

    def forward(self, x, final_relu: bool = False, out_dtype: Optional[torch.dtype] = None, final_dropout: float = 0.0):
        """Extensions (defaults = reference behaviour): `final_relu` applies the ReLU every caller on the path wraps
        around the MLP (`F.relu(self.f_enc(x))`, reference src/layers.py:631,634) and `final_dropout` the dropout that
        follows it (:632, training only), both inside the last fused pass; `out_dtype` is honoured by the fused paths."""
        if self._tc_ok(x):
            l0, l1 = self.normalizations
            y = _lib.mlp2_fwd(x.contiguous(), self.lins[0].weight, self.lins[0].bias, self.lins[1].weight,
                              self.lins[1].bias, self._ln_tuple(l0), self._ln_tuple(l1), final_relu,
                              out_dtype or torch.float32)
            return F.dropout(y, p=final_dropout, training=True) if (final_dropout > 0 and self.training) else y
        if self._chain_ok(x):
            return self._forward_chain(x, final_relu, out_dtype, final_dropout)
        if x.dtype != self.lins[0].weight.dtype:
            x = x.to(self.lins[0].weight.dtype)
        x = self.normalizations[0](x)
        for i, lin in enumerate(self.lins[:-1]):
            x = F.relu(lin(x), inplace=True)
            x = self.normalizations[i + 1](x)
            x = F.dropout(x, p=self.dropout, training=self.training)
        x = self.lins[-1](x)
        if final_relu:
            x = F.relu(x)
        return F.dropout(x, p=final_dropout, training=self.training) if final_dropout > 0 else x

    # -- chain path (training AND inference, fp32 or bf16 mode): bias-free GEMMs with ONE fused rowop pass between them
    #    (bias, ReLU, LayerNorm, dropout -- forward and backward are each one kernel, allset_rowop_fwd / _bwd).  In bf16
    #    mode the GEMMs take bf16 operands on the tensor cores and the activations between them are bf16; parameters,
    #    LayerNorm statistics, gradients of the parameters and the rows handed to the next half layer stay fp32. ------
    def compute_dtype(self) -> torch.dtype:
        return torch.bfloat16 if self.tc_dtype == torch.bfloat16 else torch.float32

    def _chain_ok(self, x) -> bool:
        if not (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16)):
            return False
        if x.shape[0] < ops.FUSED_DENSE_MIN_ROWS or self.lins[0].weight.dtype != torch.float32:
            return False
        if not all(isinstance(n, (nn.LayerNorm, nn.Identity)) for n in self.normalizations):
            return False
        return all(lin.out_features in _lib.ROWOP_WIDTHS for lin in self.lins[:-1])

    def _forward_chain(self, x, final_relu, out_dtype, final_dropout):
        cd = self.compute_dtype()
        io = x.dtype if out_dtype is None else out_dtype
        p = self.dropout if self.training else 0.0
        pf = final_dropout if self.training else 0.0
        if (not torch.is_grad_enabled() and p == 0.0 and pf == 0.0 and x.is_contiguous()
                and (cd == torch.bfloat16 or (x.dtype == torch.float32 and io == torch.float32))
                and all(ops.tc_linear_ok(x, lin.weight) for lin in self.lins)):
            # inference: every Linear is ONE tcgen05 launch with the LayerNorm in front of it, its bias and the ReLU behind
            # it inside the kernel (rows read once, written once per Linear); fp32 mode in split precision
            h = x
            for i, lin in enumerate(self.lins):
                last = i == len(self.lins) - 1
                h = ops.linear_fused(h, lin.weight, lin.bias, ln=self._ln_tuple(self.normalizations[i]),
                                     relu=final_relu if last else True, compute_dtype=cd, out_dtype=io if last else cd)
            return h
        n0 = self.normalizations[0]
        if isinstance(n0, nn.LayerNorm):
            h = ops.rowop(x, gamma=n0.weight, beta=n0.bias, eps=n0.eps, out_dtype=cd)
        else:
            h = x if x.dtype == cd else x.to(cd)
        for i, lin in enumerate(self.lins[:-1]):
            n = self.normalizations[i + 1]
            h = ops.rowop(ops.linear_nb(h, lin.weight), lin.bias, relu=True, drop_p=p, out_dtype=cd, **self._ln(n))
        last = self.lins[-1]
        return ops.rowop(ops.linear_nb(h, last.weight), last.bias, relu=final_relu, drop_p=pf, out_dtype=io)

    # -- tensor-core path: the whole two-layer MLP in one tcgen05 kernel (eval mode, bf16 operands) ---------------------
    def _tc_ok(self, x) -> bool:
        if self.tc_dtype != torch.bfloat16 or torch.is_grad_enabled() or len(self.lins) != 2:
            return False
        if self.training and self.dropout > 0:       # the fused kernel has no dropout between the Linears (MC dropout,
            return False                             # no-grad forwards in train mode): take the path that applies it
        if not (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16)):
            return False
        d = x.shape[1]
        if d not in _lib.MLP2_WIDTHS or x.shape[0] < ops.FUSED_DENSE_MIN_ROWS:
            return False
        if any(tuple(lin.weight.shape) != (d, d) or lin.weight.dtype != torch.float32 for lin in self.lins):
            return False
        return all(isinstance(n, (nn.LayerNorm, nn.Identity)) for n in self.normalizations)

    @staticmethod
    def _ln_tuple(norm):
        return (norm.weight, norm.bias, norm.eps) if isinstance(norm, nn.LayerNorm) else None

    @staticmethod
    def _ln(norm):
        if isinstance(norm, nn.LayerNorm):
            return dict(gamma=norm.weight, beta=norm.bias, eps=norm.eps)
        return {}


def _resolve(edge_index, n_src: int) -> Incidence:
    if isinstance(edge_index, Incidence) or hasattr(edge_index, 'sharded_segment_reduce'):
        return edge_index.with_n_src(n_src)
    if not isinstance(edge_index, Tensor):
        raise TypeError('edge_index must be a [2, nnz] tensor or an allset_b200.Incidence')
    if not edge_index.is_cuda:
        raise RuntimeError('allset_b200 layers run on CUDA only (no CPU fallback): edge_index is on %s'
                           % edge_index.device)
    return incidence_of(edge_index, n_src)


class PMA(nn.Module):
    """Pooling by multi-head attention with ONE learned seed per head (reference src/layers.py:42-199).

    forward: K = lin_K(x), V = lin_V(x), score = <K_h, att_r_h> per source row; per target segment a softmax of
    leaky_relu(score) weights the V rows; + seed; LN0; LN1(out + relu(rFF(out))).  The segment part (everything
    the reference does inside `propagate`, plus the seed residual) is one CUDA kernel."""

    def __init__(self, in_channels, hid_dim, out_channels, num_layers, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, bias=False, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.hidden = hid_dim // heads
        self.out_channels = out_channels
        self.heads = heads
        self.concat = concat
        self.negative_slope = negative_slope
        self.dropout = 0.          # the reference hard-wires attention dropout off (src/layers.py:63)
        self.aggr = 'add'
        self.lin_K = Linear(in_channels, self.heads * self.hidden)
        self.lin_V = Linear(in_channels, self.heads * self.hidden)
        self.att_r = Parameter(torch.Tensor(1, heads, self.hidden))      # seed
        self.rFF = MLP(in_channels=self.heads * self.hidden, hidden_channels=self.heads * self.hidden,
                       out_channels=out_channels, num_layers=num_layers, dropout=.0, Normalization='None')
        self.ln0 = nn.LayerNorm(self.heads * self.hidden)
        self.ln1 = nn.LayerNorm(self.heads * self.hidden)
        self.register_parameter('bias', None)
        self.agg_dtype: Optional[torch.dtype] = None   # storage dtype of the gathered rows (None = x.dtype)
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.lin_K.weight)       # biases keep their construction-time values, as in the reference
        glorot(self.lin_V.weight)
        self.rFF.reset_parameters()
        self.ln0.reset_parameters()
        self.ln1.reset_parameters()
        nn.init.xavier_uniform_(self.att_r)

    def forward(self, x, edge_index, size=None, return_attention_weights=None, relu_out: bool = False,
                out_dropout: float = 0.0):
        """Extensions (defaults = reference behaviour): `relu_out` also applies the ReLU SetGNN.forward wraps around every
        half layer (reference src/models.py:475,478) and `out_dropout` the dropout that follows it (:476,479, training
        only), inside the last fused pass where there is one."""
        assert x.dim() == 2, 'Static graphs not supported in `GATConv`.'
        H, C = self.heads, self.hidden
        inc = _resolve(edge_index, x.size(0))
        # score = (lin_K(x).view(-1,H,C) * att_r).sum(-1)  (reference :128,:130), folded: there is one seed per head,
        # so the score is linear in x with W_eff[h,:] = sum_c att_r[h,c] W_K[hC+c,:] -- an [n_src,in]x[in,H] GEMV instead
        # of the [n_src,in]x[in,H*C] GEMM + multiply + reduce.  Same algebra, differentiable, no 1/sqrt(C).
        seed = self.att_r.view(H, C)
        w_eff = (self.lin_K.weight.view(H, C, -1) * seed.unsqueeze(-1)).sum(dim=1)         # [H, in]
        b_eff = (self.lin_K.bias.view(H, C) * seed).sum(dim=1)                             # [H]
        want_alpha = isinstance(return_attention_weights, bool)
        pdrop = out_dropout if self.training else 0.0
        out = None
        tc_v = self._tc_v_ok(x)
        tc_score = tc_v and H * H * C <= 1024               # w_eff must fit the kernel's 4 KB side buffer
        chain = self._chain_ok(x)
        cd = self.rFF.compute_dtype()
        # rows handed to the next half layer: fp32, except while training in bf16 mode (bf16 activations end to end)
        out_dt = torch.bfloat16 if (chain and cd == torch.bfloat16 and torch.is_grad_enabled()) else w_eff.dtype
        if tc_score:
            score = None
        elif chain and cd == torch.bfloat16:
            # bf16 mode at scale: the skinny score GEMM takes the same bf16 rows as lin_V (fp32 accumulate AND fp32 result)
            xb = x if x.dtype == cd else x.to(cd)
            score = ops.linear_nb(xb, w_eff, out_fp32=True) + b_eff                         # [n_src, H] fp32
        else:
            xf = x if x.dtype == w_eff.dtype else x.to(w_eff.dtype)
            score = F.linear(xf, w_eff, b_eff)                                              # [n_src, H] fp32
        if tc_v:
            # bf16 mode: V = lin_V(x) as ONE tcgen05 kernel that writes the bf16 rows the aggregation gathers; the same
            # launch computes the fp32 scores in its producer warps (no second pass over x)
            xc = x.contiguous()
            w_eff_c, b_eff_c = w_eff.contiguous(), b_eff.contiguous()

            def lin_v(out_view=None):
                if tc_score:
                    return _lib.linear_score_fwd(xc, self.lin_V.weight, self.lin_V.bias, w_eff_c, b_eff_c,
                                                 out_dtype=torch.bfloat16, out=out_view)
                return _lib.mlp2_fwd(xc, self.lin_V.weight, self.lin_V.bias, None, None, None, None, False,
                                     torch.bfloat16, out=out_view), score

            if not want_alpha and self._packed_ok(xc, inc):
                # scores too large for L2: one packed [values | scores] record per source row, so that a gather touches
                # one contiguous record per incidence instead of a row plus a 32-byte score in another 128-byte line
                _, v, s = _lib.packed_pma_records(xc.shape[0], H * C, H, torch.bfloat16, xc.device)
                _, score = lin_v(v)
                s.copy_(score)
                t = inc.by_tgt
                try:
                    out, _ = _lib.pma_fwd_strided(v, s, self.att_r.detach().float().reshape(-1).contiguous(), H, C,
                                                  self.negative_slope, t.rowptr, t.col, t.n_tgt)
                    alpha = None
                except _lib.Unsupported:
                    v = v.contiguous()
            else:
                v, score = lin_v()
        elif chain:
            # training / fp32 mode at scale: bias-free GEMM (bf16 operands in bf16 mode) + one rowop pass that adds the
            # bias and writes the rows in the storage dtype the aggregation gathers
            xb = x if x.dtype == cd else x.to(cd)
            v = ops.linear_bias_act(xb, self.lin_V.weight, self.lin_V.bias, out_dtype=self.agg_dtype or cd)
        else:
            x_V = self.lin_V(x if x.dtype == w_eff.dtype else x.to(w_eff.dtype))
            v = x_V if self.agg_dtype is None else x_V.to(self.agg_dtype)
        if out is None:
            out, alpha = ops.pma_aggregate(v, score, self.att_r, inc, H, self.negative_slope, return_alpha=want_alpha)
        applied = False
        if self.rFF._tc_ok(out):
            # bf16 mode: ln0 -> rFF -> ln1(residual + relu(.)) [-> the caller's ReLU] as ONE tcgen05 kernel reading the
            # aggregated rows in their storage dtype
            l0, l1 = self.rFF.lins
            out = _lib.pma_tail_fwd(out.contiguous(), (self.ln0.weight, self.ln0.bias, self.ln0.eps), l0.weight, l0.bias,
                                    l1.weight, l1.bias, (self.ln1.weight, self.ln1.bias, self.ln1.eps),
                                    relu_final=relu_out, out_dtype=score.dtype)
            if pdrop > 0:
                out = F.dropout(out, p=pdrop, training=True)
            applied = True
        elif chain and self.rFF._chain_ok(out):
            y = ops.rowop(out, gamma=self.ln0.weight, beta=self.ln0.bias, eps=self.ln0.eps, out_dtype=cd)   # ln0 (:155)
            a = y
            for lin in self.rFF.lins[:-1]:                                   # rFF: no norms, no dropout (:76-80)
                a = ops.linear_bias_act(a, lin.weight, lin.bias, relu=True, out_dtype=cd)
            last = self.rFF.lins[-1]
            out = ops.rowop(ops.linear_nb(a, last.weight), last.bias, relu=True, residual=y, gamma=self.ln1.weight,
                            beta=self.ln1.bias, eps=self.ln1.eps, relu_out=relu_out, drop_p=pdrop,
                            out_dtype=out_dt)                                # ln1(y + relu(rFF(y))) [relu, dropout], one pass
            applied = True
        else:
            out = self.ln0(out.to(score.dtype))
            out = self.ln1(out + F.relu(self.rFF(out)))
        if not applied:
            if relu_out:
                out = F.relu(out)
            if pdrop > 0:
                out = F.dropout(out, p=pdrop, training=True)
        if want_alpha:
            return out, (edge_index, alpha)
        return out

    def _chain_ok(self, x) -> bool:
        d = self.heads * self.hidden
        return (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16)
                and x.shape[0] >= ops.FUSED_DENSE_MIN_ROWS and d in _lib.ROWOP_WIDTHS
                and self.lin_V.weight.dtype == torch.float32)

    # Packed [values | scores] records for the gather (allset_pma_fwd_strided) are OFF by default: measured on the
    # 10 M-vertex graph they are slower (4.03 vs 3.85 ms V->E) -- HBM fills L2 in 64-byte units, so a 288-byte record
    # costs the same 320 bytes as a 256-byte row plus a 32-byte score elsewhere, and it straddles three 128-byte lines.
    # Set to a byte count (e.g. 96 << 20 = "scores do not fit L2") to enable.
    PACKED_MIN_SCORE_BYTES = None

    def _packed_ok(self, x, inc) -> bool:
        H, d = self.heads, self.heads * self.hidden
        if not isinstance(inc, Incidence):
            return False
        t = inc.by_tgt
        if self.PACKED_MIN_SCORE_BYTES is None or x.shape[0] * H * 4 < self.PACKED_MIN_SCORE_BYTES or H % 4 != 0:
            return False
        if t.long_ids is not None and t.long_ids.numel() > 0:
            return False
        return bool(_lib.lib().allset_stream_eligible(_lib.BF16, d, t.n_tgt))

    def _tc_v_ok(self, x) -> bool:
        """lin_V on the tensor cores: bf16 mode, eval, square Linear of a width the tcgen05 kernel has."""
        d = self.heads * self.hidden
        return (self.agg_dtype == torch.bfloat16 and not torch.is_grad_enabled() and x.is_cuda and x.dim() == 2
                and x.dtype in (torch.float32, torch.bfloat16) and x.shape[1] == d and d in _lib.MLP2_WIDTHS
                and x.shape[0] >= ops.FUSED_DENSE_MIN_ROWS and self.lin_V.weight.dtype == torch.float32)

    def __repr__(self):
        return '{}({}, {}, heads={})'.format(self.__class__.__name__, self.in_channels, self.out_channels, self.heads)


class HalfNLHconv(nn.Module):
    """One half layer, V->E or E->V (reference src/layers.py:582-656).

    attention=True : PMA.   attention=False : relu(f_enc(x)) -> dropout -> segmented sum/mean of norm_e * x[src_e]
    -> relu(f_dec(.)).  `edge_index` row 0 = source rows, row 1 = target rows; output rows = max(target)+1."""

    def __init__(self, in_dim, hid_dim, out_dim, num_layers, dropout, Normalization='bn', InputNorm=False,
                 heads=1, attention=True):
        super().__init__()
        self.attention = attention
        self.dropout = dropout
        self.agg_dtype: Optional[torch.dtype] = None
        if self.attention:
            self.prop = PMA(in_dim, hid_dim, out_dim, num_layers, heads=heads)
        elif num_layers > 0:
            self.f_enc = MLP(in_dim, hid_dim, hid_dim, num_layers, dropout, Normalization, InputNorm)
            self.f_dec = MLP(hid_dim, hid_dim, out_dim, num_layers, dropout, Normalization, InputNorm)
        else:
            self.f_enc = nn.Identity()
            self.f_dec = nn.Identity()

    def reset_parameters(self):
        if self.attention:
            self.prop.reset_parameters()
        else:
            for f in (self.f_enc, self.f_dec):
                if not isinstance(f, nn.Identity):
                    f.reset_parameters()

    def set_agg_dtype(self, dtype: Optional[torch.dtype]):
        """bf16 storage for the gathered rows also switches the eval-mode MLPs around the aggregation to the tcgen05
        kernel with bf16 operands (same 1e-2 accuracy class); fp32 / None keeps fp32 numerics end to end."""
        self.agg_dtype = dtype
        if self.attention:
            self.prop.agg_dtype = dtype
            self.prop.rFF.tc_dtype = dtype if dtype == torch.bfloat16 else None
        else:
            for f in (self.f_enc, self.f_dec):
                if isinstance(f, MLP):
                    f.tc_dtype = dtype if dtype == torch.bfloat16 else None

    def forward(self, x, edge_index, norm, aggr='add', relu_out: bool = False, out_dropout: float = 0.0):
        """Extensions: `relu_out` folds SetGNN.forward's `F.relu(conv(.))` (reference src/models.py:475,478) into the layer
        -- the identity for a non-attention layer, which already ends in relu(f_dec(.)) -- and `out_dropout` the dropout
        SetGNN applies to the result (:476,479; training only)."""
        if self.attention:
            return self.prop(x, edge_index, relu_out=relu_out, out_dropout=out_dropout)   # norm, aggr ignored (reference)
        if aggr is None:
            raise ValueError('aggr was not passed!')
        io_dtype = x.dtype
        if isinstance(self.f_enc, MLP):
            # relu(f_enc(x)) and the dropout behind it (:631-632) end in the storage dtype the aggregation gathers
            x = self.f_enc(x, final_relu=True, out_dtype=self.agg_dtype, final_dropout=self.dropout)
        else:
            x = F.dropout(F.relu(self.f_enc(x)), p=self.dropout, training=self.training)
        inc = _resolve(edge_index, x.size(0))
        weight = None
        if norm is not None and not (not norm.requires_grad and inc.weights_all_one(norm)):
            weight = norm
        xs = x if self.agg_dtype is None else x.to(self.agg_dtype)
        x = ops.segment_reduce(xs, inc, weight, aggr)
        if isinstance(self.f_dec, MLP):
            if not (self.f_dec._tc_ok(x) or self.f_dec._chain_ok(x)):     # the fused paths read the storage dtype directly
                x = x.to(io_dtype)
            # rows handed to the next half layer: the caller's dtype, except while training in bf16 mode (bf16 activations
            # end to end: the next f_enc reads half the bytes)
            out_dt = torch.bfloat16 if (self.agg_dtype == torch.bfloat16 and torch.is_grad_enabled()
                                        and self.f_dec._chain_ok(x)) else io_dtype
            x = self.f_dec(x, final_relu=True, out_dtype=out_dt, final_dropout=out_dropout)
        else:
            x = F.relu(self.f_dec(x.to(io_dtype)))
            if out_dropout > 0:
                x = F.dropout(x, p=out_dropout, training=self.training)
        return x
