"""CPU oracle for the AllSet V->E / E->V multiset-aggregation path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it.  `allset_b200/` never does, and raises if its CUDA library is
missing instead of falling back to anything in here.

What it is: a functional (state_dict-driven) restatement, in plain torch CPU ops, of
  * the reference's own Python for this path  -- `/root/reference/src/layers.py` (PMA :42-199, MLP :496-579,
    HalfNLHconv :582-656) and `/root/reference/src/models.py` (SetGNN :295-484), and
  * the three third-party primitives that path calls and that are NOT vendored under `/root/reference`:
    torch-scatter 2.0.4 `scatter` (sum/mean/max), torch-geometric 1.6.3 `utils.softmax` and
    `MessagePassing.propagate` (pins: reference README.md:18-22; published semantics restated in SURVEY.md
    Appendix C).

Parity pinning: the reference ships no tests or golden vectors for this path, so the oracle
is pinned against outputs OF THE REFERENCE ITSELF: `oracle/make_golden.py` imports the reference's unmodified
`layers.py` / `models.py` / `preprocessing.py` in the dev container (third-party imports satisfied by
`oracle/shims/`), runs seeded configurations and commits inputs, state_dicts, outputs, intermediates and gradients
under `tests/golden/`.  `tests/test_oracle_golden.py` checks every function below against those files.
What stays unpinned: torch-scatter / PyG themselves are restated from their documented behaviour, not executed.

Every function takes `params`: a dict with exactly the reference's `state_dict()` keys (e.g.
`V2EConvs.0.f_enc.lins.0.weight`, `V2EConvs.0.prop.att_r`), so weights move freely between the reference,
the oracle and `allset_b200`.  Works in fp32 or fp64 (dtype follows `x`/`params`), and is differentiable, so it is
also the gradient oracle.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------------------
# third-party primitives (torch-scatter 2.0.4 / torch-geometric 1.6.3), restated
# --------------------------------------------------------------------------------------------------------------
def implied_rows(index: Tensor) -> int:
    """torch_scatter sizes its output as index.max()+1 when no dim_size is given (0 rows for an empty index).

    The reference deliberately drops dim_size in `aggregate` (layers.py:179-194, 641-656), so this rule decides
    the row count of every V->E / E->V result."""
    return 0 if index.numel() == 0 else int(index.max()) + 1


def scatter_rows(src: Tensor, index: Tensor, reduce: str = 'sum', rows: Optional[int] = None) -> Tensor:
    """`torch_scatter.scatter(src, index, dim=0 (node axis), reduce=...)` for src [nnz, ...], index [nnz].

    sum/add: zeros(rows).scatter_add_ ; mean: sum / clamp(count, 1) ; max/min: empty segments stay 0."""
    rows = implied_rows(index) if rows is None else rows
    out = torch.zeros((rows,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if src.numel() == 0:
        return out
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    if reduce in ('sum', 'add'):
        return out.scatter_add_(0, idx, src)
    if reduce == 'mean':
        out = out.scatter_add_(0, idx, src)
        cnt = torch.zeros(rows, dtype=src.dtype, device=src.device)
        cnt.scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1)
        return out / cnt.view((-1,) + (1,) * (src.dim() - 1))
    if reduce in ('max', 'min'):
        return out.scatter_reduce(0, idx, src, reduce='a' + reduce, include_self=False)
    raise ValueError('unknown reduce %r' % (reduce,))


def segment_softmax(score: Tensor, index: Tensor, rows: Optional[int] = None) -> Tensor:
    """PyG 1.6.3 `softmax(src, index, ptr=None, num_nodes)`: subtract the per-segment max, exp, divide by
    (segment sum + 1e-16).  score [nnz, H], index [nnz]."""
    rows = implied_rows(index) if rows is None else rows
    seg_max = scatter_rows(score, index, 'max', rows)
    e = (score - seg_max[index]).exp()
    seg_sum = scatter_rows(e, index, 'sum', rows)
    return e / (seg_sum[index] + 1e-16)


def gather_rows(x: Tensor, index: Tensor) -> Tensor:
    """`MessagePassing.__lift__`: x.index_select(node_dim, edge_index[0]) (flow = source_to_target)."""
    return x.index_select(0, index)


# --------------------------------------------------------------------------------------------------------------
# the three raw aggregation ops the CUDA kernels replace
# --------------------------------------------------------------------------------------------------------------
def aggregate_sum_mean(x: Tensor, src: Tensor, tgt: Tensor, norm: Optional[Tensor], aggr: str) -> Tensor:
    """HalfNLHconv propagate (layers.py:633, message :638-639, aggregate :641-656):
    out[t] = reduce_{e: tgt[e]=t} norm[e] * x[src[e]],  rows = tgt.max()+1."""
    msg = gather_rows(x, src)
    if norm is not None:
        msg = norm.view(-1, 1) * msg          # int64 ones promote to the feature dtype (preprocessing.py:454)
    return scatter_rows(msg, tgt, aggr)


def aggregate_pma(v: Tensor, score: Tensor, seed: Tensor, src: Tensor, tgt: Tensor,
                  negative_slope: float = 0.2) -> Tuple[Tensor, Tensor]:
    """PMA propagate (layers.py:145-146, message :168-177, aggregate :179-194) plus the seed residual (:153).

    v [n_src, H, C], score [n_src, H] (= alpha_r, :130), seed [1, H, C] (= att_r).
    Returns (out [rows, H, C] with the seed added, alpha [nnz, H])."""
    a = F.leaky_relu(gather_rows(score, src), negative_slope)
    alpha = segment_softmax(a, tgt)
    out = scatter_rows(gather_rows(v, src) * alpha.unsqueeze(-1), tgt, 'sum')
    return out + seed, alpha


# --------------------------------------------------------------------------------------------------------------
# reference modules, restated functionally over the reference's state_dict keys
# --------------------------------------------------------------------------------------------------------------
def _count(params: Dict[str, Tensor], prefix: str, what: str) -> int:
    n = 0
    while (prefix + '%s.%d.weight' % (what, n)) in params:
        n += 1
    return n


def _norm_layer(params, key: str, x: Tensor, training: bool) -> Tensor:
    """normalizations[i] of MLP (layers.py:499-562): LayerNorm, BatchNorm1d or Identity -- detected from the keys."""
    w = params.get(key + '.weight')
    if w is None:
        return x                                                   # nn.Identity has no parameters
    b = params[key + '.bias']
    if (key + '.running_mean') in params:                          # BatchNorm1d
        return F.batch_norm(x, params[key + '.running_mean'].clone(), params[key + '.running_var'].clone(),
                            w, b, training=training, momentum=0.1, eps=1e-5)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def mlp(params, prefix: str, x: Tensor, dropout: float = 0.0, training: bool = False) -> Tensor:
    """MLP.forward (layers.py:571-579): norm0 -> [Linear -> ReLU -> norm -> dropout] x (L-1) -> Linear."""
    n_lin = _count(params, prefix, 'lins')
    if n_lin == 0:
        return x                                                   # nn.Identity f_enc/f_dec (layers.py:605-607)
    x = _norm_layer(params, prefix + 'normalizations.0', x, training)
    for i in range(n_lin - 1):
        x = F.linear(x, params[prefix + 'lins.%d.weight' % i], params[prefix + 'lins.%d.bias' % i])
        x = F.relu(x)
        x = _norm_layer(params, prefix + 'normalizations.%d' % (i + 1), x, training)
        x = F.dropout(x, p=dropout, training=training)
    i = n_lin - 1
    return F.linear(x, params[prefix + 'lins.%d.weight' % i], params[prefix + 'lins.%d.bias' % i])


def pma(params, prefix: str, x: Tensor, src: Tensor, tgt: Tensor, heads: int,
        negative_slope: float = 0.2) -> Tuple[Tensor, Tensor]:
    """PMA.forward (layers.py:107-166).  Returns (out [rows, H*C], alpha [nnz, H])."""
    wk, bk = params[prefix + 'lin_K.weight'], params[prefix + 'lin_K.bias']
    wv, bv = params[prefix + 'lin_V.weight'], params[prefix + 'lin_V.bias']
    seed = params[prefix + 'att_r']                                # [1, H, C]
    H, C = heads, wk.shape[0] // heads
    x_k = F.linear(x, wk, bk).view(-1, H, C)                       # :128
    x_v = F.linear(x, wv, bv).view(-1, H, C)                       # :129
    score = (x_k * seed).sum(dim=-1)                               # :130  (no 1/sqrt(C))
    out, alpha = aggregate_pma(x_v, score, seed, src, tgt, negative_slope)   # :145-153
    out = out.reshape(-1, H * C)
    out = F.layer_norm(out, (H * C,), params[prefix + 'ln0.weight'], params[prefix + 'ln0.bias'], 1e-5)   # :155
    ff = mlp(params, prefix + 'rFF.', out)                         # rFF: Normalization='None', dropout 0 (:76-80)
    out = F.layer_norm(out + F.relu(ff), (H * C,), params[prefix + 'ln1.weight'], params[prefix + 'ln1.bias'], 1e-5)
    return out, alpha                                              # :157


def half_nlh_conv(params, prefix: str, x: Tensor, src: Tensor, tgt: Tensor, norm: Optional[Tensor], aggr: str,
                  attention: bool, heads: int = 1, dropout: float = 0.0, training: bool = False) -> Tensor:
    """HalfNLHconv.forward (layers.py:623-636): one half layer, source rows `x` -> target rows."""
    if attention:
        return pma(params, prefix + 'prop.', x, src, tgt, heads)[0]            # :628-629 (norm, aggr ignored)
    x = F.relu(mlp(params, prefix + 'f_enc.', x, dropout, training))            # :631
    x = F.dropout(x, p=dropout, training=training)                              # :632
    x = aggregate_sum_mean(x, src, tgt, norm, aggr)                             # :633
    return F.relu(mlp(params, prefix + 'f_dec.', x, dropout, training))         # :634


def setgnn(params, x: Tensor, edge_index: Tensor, norm: Optional[Tensor], *, PMA: bool, heads: int = 1,
           aggregate: str = 'mean', dropout: float = 0.0, GPR: bool = False, LearnMask: bool = False,
           training: bool = False, input_dropout: float = 0.2) -> Tuple[Tensor, List[Tensor]]:
    """SetGNN.forward (models.py:435-484).  Returns (logits, [V2E_0, E2V_0, V2E_1, ...] post-ReLU intermediates).

    `edge_index` is NOT modified here (the reference zero-bases row 1 in place, models.py:453-454; the product
    reproduces that side effect, the oracle just computes with the shifted ids)."""
    node = edge_index[0]
    he = edge_index[1] - edge_index[1].min()                                    # :453-454
    if LearnMask:
        norm = params['Importance'] * norm                                      # :451-452
    n_layers = _count_layers(params)
    taps: List[Tensor] = []
    kw = dict(aggr=aggregate, attention=PMA, heads=heads, dropout=dropout, training=training)
    if n_layers == 0:
        return mlp(params, 'classifier.', x, dropout, training), taps           # models.py:339-346: classifier only
    if GPR:                                                                      # :457-471
        xs = [F.relu(mlp(params, 'MLP.', x, dropout, training))]
        for i in range(n_layers):
            x = F.relu(half_nlh_conv(params, 'V2EConvs.%d.' % i, x, node, he, norm, **kw))
            taps.append(x)
            x = F.dropout(x, p=dropout, training=training)
            x = F.relu(half_nlh_conv(params, 'E2VConvs.%d.' % i, x, he, node, norm, **kw))
            taps.append(x)
            xs.append(x)
            x = F.dropout(x, p=dropout, training=training)
        x = torch.stack(xs, dim=-1)
        x = F.linear(x, params['GPRweights.weight']).squeeze()
        return mlp(params, 'classifier.', x, dropout, training), taps
    x = F.dropout(x, p=input_dropout, training=training)                        # :473
    for i in range(n_layers):                                                   # :474-481
        x = F.relu(half_nlh_conv(params, 'V2EConvs.%d.' % i, x, node, he, norm, **kw))
        taps.append(x)
        x = F.dropout(x, p=dropout, training=training)
        x = F.relu(half_nlh_conv(params, 'E2VConvs.%d.' % i, x, he, node, norm, **kw))
        taps.append(x)
        x = F.dropout(x, p=dropout, training=training)
    return mlp(params, 'classifier.', x, dropout, training), taps               # :482


def _count_layers(params) -> int:
    # bnV2Es.i exists for every layer (models.py:357,378) even when the convs themselves hold no parameters
    # (MLP_num_layers == 0 -> Identity f_enc / f_dec, layers.py:605-607)
    n = 0
    while any(k.startswith('V2EConvs.%d.' % n) or k.startswith('bnV2Es.%d.' % n) for k in params):
        n += 1
    return n


# --------------------------------------------------------------------------------------------------------------
# the reference's CPU op sequence for ONE V->E + E->V pair, used by bench.py's cpu_baseline / --impl reference
# --------------------------------------------------------------------------------------------------------------
def layer_pair_sum(x_v: Tensor, node: Tensor, he: Tensor, norm: Optional[Tensor], aggr: str = 'sum') -> Tuple[Tensor, Tensor]:
    """index_select -> norm * x_j -> scatter_add_ twice (V->E then E->V), exactly the op triple the reference's
    HalfNLHconv launches per direction, without the dense MLPs."""
    x_e = aggregate_sum_mean(x_v, node, he, norm, aggr)
    x_v2 = aggregate_sum_mean(x_e, he, node, norm, aggr)
    return x_e, x_v2


def layer_pair_pma(v_v: Tensor, score_v: Tensor, v_e_fn, seed: Tensor, node: Tensor, he: Tensor):
    """Two PMA aggregations (V->E then E->V) on raw value/score tensors; `v_e_fn(out_e)` produces the
    (value, score) pair of the hyperedge side from the V->E result."""
    out_e, _ = aggregate_pma(v_v, score_v, seed, node, he)
    v_e, score_e = v_e_fn(out_e)
    out_v, _ = aggregate_pma(v_e, score_e, seed, he, node)
    return out_e, out_v


def config_namespace(**kw) -> SimpleNamespace:
    """The `args` namespace SetGNN reads (train.py:221-289 defaults; models.py:321-327,340-412)."""
    d = dict(All_num_layers=2, dropout=0.5, aggregate='mean', normalization='ln', deepset_input_norm=True,
             GPR=False, LearnMask=False, num_features=None, MLP_hidden=64, MLP_num_layers=2, heads=1, PMA=True,
             Classifier_hidden=64, Classifier_num_layers=2, num_classes=None)
    d.update(kw)
    return SimpleNamespace(**d)


# ----------------------------------------------------------------------------------------------------------------------
# Split-precision Linear (test infrastructure for allset_linear_fwd / allset_linear_wgrad, csrc/mlp_tcgen05.cuh MODE 3 and
# csrc/linear_wgrad.cuh): the ARITHMETIC the tensor-core kernels perform, restated on the CPU.  The reference computes
# `x @ W.t()` in fp32 (nn.Linear, reference src/layers.py:575); the kernels cut both operands into bf16 terms, whose
# pairwise products are exact in fp32, and accumulate the products in fp32.
# ----------------------------------------------------------------------------------------------------------------------
def bf16_terms(x: Tensor, n_terms: int) -> List[Tensor]:
    """x (fp32) as n_terms bf16-representable fp32 tensors: t0 = bf16(x), t1 = bf16(x - t0), ... (round to nearest even)."""
    terms, rest = [], x.float()
    for _ in range(n_terms):
        t = rest.to(torch.bfloat16).float()
        terms.append(t)
        rest = rest - t
    return terms


def linear_split(x: Tensor, w: Tensor, n_terms: int = 3) -> Tensor:
    """x @ w.t() from bf16 terms: all products x_i w_j with i + j < n_terms (n_terms = 3: the six products of
    allset_linear_fwd; n_terms = 2: the three of the usual "3x" split), smallest first, fp32 accumulation."""
    xs, ws = bf16_terms(x, n_terms), bf16_terms(w, n_terms)
    pairs = sorted(((i, j) for i in range(n_terms) for j in range(n_terms) if i + j < n_terms), key=lambda p: -(p[0] + p[1]))
    acc = torch.zeros(x.shape[0], w.shape[0], dtype=torch.float32)
    for i, j in pairs:
        acc = acc + xs[i] @ ws[j].t()
    return acc
