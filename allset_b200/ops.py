"""Differentiable operators over an `Incidence`: the two aggregations of the AllSet layer.

    segment_reduce  == propagate of HalfNLHconv       (reference src/layers.py:633, message :638-639, aggregate :641-656)
    pma_aggregate   == propagate of PMA + seed add    (reference src/layers.py:145-153, message :168-177, aggregate :179-194)

Forward and backward are each ONE launch of a hand-written sm_100a kernel through the C ABI (allset_b200/_lib.py);
the backward w.r.t. the gathered rows is the same gather kernel on the transposed CSR, so nothing here uses
atomics and results are run-to-run deterministic.  No CPU path: non-CUDA tensors raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .graph import Incidence


def _storage(x: torch.Tensor) -> torch.Tensor:
    if x.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError('aggregation kernels store rows as float32 or bfloat16, got %s' % x.dtype)
    return x.contiguous()


class _SegReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, inc: Incidence, mean: bool):
        t = inc.by_tgt
        x = _storage(x)
        w_csr = None
        if weight is not None:
            w_csr = weight.detach().float().index_select(0, t.perm64)
        out = _lib.segreduce_fwd(x, t.rowptr, t.col, t.n_tgt, mean, w=w_csr, long_ids=t.long_ids,
                                 long_threshold=t.long_threshold, max_segment_len=t.max_len)
        ctx.inc, ctx.mean = inc, mean
        ctx.has_weight = weight is not None
        ctx.weight_dtype = None if weight is None else weight.dtype
        ctx.save_for_backward(x if (weight is not None and ctx.needs_input_grad[1]) else None,
                              weight.detach() if weight is not None else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        inc, mean = ctx.inc, ctx.mean
        x, weight = ctx.saved_tensors
        t, s = inc.by_tgt, inc.by_src
        grad_out = _storage(grad_out)
        grad_x = grad_w = None
        if ctx.needs_input_grad[0]:
            # d out[t] / d x[s] = w_e (/ count_t): the same gather-reduce over the transposed CSR
            w_T = None if weight is None else weight.float().index_select(0, s.perm64)
            grad_x = _lib.segreduce_fwd(grad_out, s.rowptr, s.col, s.n_tgt, False, w=w_T,
                                        src_scale=t.inv_count if mean else None,
                                        long_ids=s.long_ids, long_threshold=s.long_threshold, max_segment_len=s.max_len)
        if ctx.has_weight and ctx.needs_input_grad[1]:
            gw_csr = _lib.segreduce_bwd_w(x, grad_out, t.rowptr, t.col, t.n_tgt,
                                          tgt_scale=t.inv_count if mean else None)
            grad_w = torch.empty_like(gw_csr)
            grad_w.index_copy_(0, t.perm64, gw_csr)           # CSR slot k came from COO position perm[k]
            grad_w = grad_w.to(ctx.weight_dtype)
        return grad_x, grad_w, None, None


def segment_reduce(x: torch.Tensor, inc: Incidence, weight: Optional[torch.Tensor] = None,
                   reduce: str = 'sum') -> torch.Tensor:
    """out[t] = reduce_{e: tgt(e)=t} weight[e] * x[src(e)]  with out.shape[0] == inc.n_tgt.

    x [n_src, d] float32 | bfloat16 (CUDA); weight [nnz] in the caller's COO order or None (= all ones);
    reduce in {'sum', 'add', 'mean'} (torch_scatter spellings reachable from reference src/train.py:38,245)."""
    if reduce not in ('sum', 'add', 'mean'):
        raise ValueError("reduce must be 'sum', 'add' or 'mean', got %r" % (reduce,))
    if hasattr(inc, 'sharded_segment_reduce'):              # one rank's share of a partitioned graph (sharded_model.py)
        if x.dim() != 2 or x.shape[0] != inc.n_src:
            raise ValueError('x must hold the %d source rows this rank owns, got %s' % (inc.n_src, tuple(x.shape)))
        return inc.sharded_segment_reduce(_storage(x), weight, reduce)
    if not x.is_cuda:
        raise RuntimeError('allset_b200.segment_reduce: CUDA tensors only (no CPU fallback); got %s' % x.device)
    if x.dim() != 2 or x.shape[0] != inc.n_src:
        raise ValueError('x must be [n_src=%d, d], got %s' % (inc.n_src, tuple(x.shape)))
    if weight is not None and weight.numel() != inc.nnz:
        raise ValueError('weight must have one entry per incidence (%d), got %d' % (inc.nnz, weight.numel()))
    return _SegReduce.apply(x, weight, inc, reduce == 'mean')


class _PMA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, score, seed, inc: Incidence, H: int, C: int, slope: float):
        t = inc.by_tgt
        v = _storage(v)
        score = score.float().contiguous()
        seed_f = seed.detach().float().reshape(-1).contiguous()
        out, stats = _lib.pma_fwd(v, score, seed_f, H, C, slope, t.rowptr, t.col, t.n_tgt, want_stats=True,
                                  long_ids=t.long_ids, long_threshold=t.long_threshold, max_segment_len=t.max_len)
        ctx.inc, ctx.H, ctx.C, ctx.slope = inc, H, C, slope
        ctx.seed_shape, ctx.seed_dtype, ctx.score_dtype = seed.shape, seed.dtype, score.dtype
        ctx.save_for_backward(v, score, seed_f, out, stats)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, grad_out, _grad_stats):
        v, score, seed_f, out, stats = ctx.saved_tensors
        inc, H, C, slope = ctx.inc, ctx.H, ctx.C, ctx.slope
        s = inc.by_src
        grad_out = _storage(grad_out)
        grad_v = grad_score = grad_seed = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # softmax backward needs D[t,h] = sum_k alpha_k <g_t, v_k> = <g_t, out_t - seed>
            D = _lib.rowdot_heads(grad_out, out, seed_f, H, C)
            grad_v, grad_score = _lib.pma_bwd(grad_out, v, score, stats, D, H, C, slope, s.rowptr, s.col, s.n_tgt,
                                              long_ids=s.long_ids, long_threshold=s.long_threshold)
        if ctx.needs_input_grad[2]:
            if grad_out.dtype == torch.bfloat16:
                # column sums of a bf16 [n_tgt, d] matrix on the tensor cores (fp32 accumulate) instead of a cast + reduce
                ones = torch.ones((1, grad_out.shape[0]), dtype=torch.bfloat16, device=grad_out.device)
                col_sum = torch.mm(ones, grad_out, out_dtype=torch.float32)
            else:
                col_sum = grad_out.sum(dim=0)
            grad_seed = col_sum.reshape(ctx.seed_shape).to(ctx.seed_dtype)
        return grad_v, grad_score, grad_seed, None, None, None, None


def pma_aggregate(v: torch.Tensor, score: torch.Tensor, seed: torch.Tensor, inc: Incidence, heads: int,
                  negative_slope: float = 0.2, return_alpha: bool = False
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Per target segment t and head h:  out[t,h,:] = sum_e softmax_t(leaky_relu(score[src(e),h])) * v[src(e),h,:] + seed[h,:]

    v [n_src, H*C] (or [n_src, H, C]); score [n_src, H]; seed [1, H, C] (PMA.att_r).  Returns (out [n_tgt, H*C],
    alpha [nnz, H] in the caller's COO order if `return_alpha` else None)."""
    if not v.is_cuda:
        raise RuntimeError('allset_b200.pma_aggregate: CUDA tensors only (no CPU fallback); got %s' % v.device)
    v2 = v.reshape(v.shape[0], -1)
    d = v2.shape[1]
    if d % heads != 0:
        raise ValueError('feature width %d is not divisible by heads=%d' % (d, heads))
    C = d // heads
    if hasattr(inc, 'sharded_pma_aggregate'):
        if return_alpha:
            raise NotImplementedError('attention weights are not assembled across ranks')
        if v2.shape[0] != inc.n_src or tuple(score.shape) != (inc.n_src, heads) or seed.numel() != d:
            raise ValueError('pma_aggregate: shape mismatch with the %d source rows this rank owns' % inc.n_src)
        return inc.sharded_pma_aggregate(_storage(v2), score, seed, heads, negative_slope), None
    if v2.shape[0] != inc.n_src or tuple(score.shape) != (inc.n_src, heads) or seed.numel() != d:
        raise ValueError('pma_aggregate: shape mismatch v%s score%s seed%s n_src=%d'
                         % (tuple(v.shape), tuple(score.shape), tuple(seed.shape), inc.n_src))
    out, stats = _PMA.apply(v2, score.float(), seed, inc, heads, C, float(negative_slope))
    alpha = None
    if return_alpha:
        t = inc.by_tgt
        a_csr = _lib.pma_alpha(score.detach().float().contiguous(), stats, heads, negative_slope, t.rowptr, t.col, t.n_tgt)
        alpha = torch.empty_like(a_csr)
        alpha.index_copy_(0, t.perm64, a_csr)
    return out, alpha


class _BiasActNorm(torch.autograd.Function):
    """out = LayerNorm(residual + relu(x + bias)) with a fused forward AND backward (allset_bias_act_norm[_bwd])."""

    @staticmethod
    def forward(ctx, x, bias, residual, gamma, beta, relu: bool, eps: float):
        x = x.contiguous()
        res = None if residual is None else residual.contiguous()
        out, stats = _lib.bias_act_norm(x, bias, relu, res, gamma, beta, eps, want_stats=True)
        ctx.relu = relu
        ctx.has = (bias is not None, residual is not None, gamma is not None, beta is not None)
        ctx.save_for_backward(x, bias, res, gamma, stats)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, bias, res, gamma, stats = ctx.saved_tensors
        has_bias, has_res, has_gamma, has_beta = ctx.has
        dx, dres, dgamma, dbeta, dbias = _lib.bias_act_norm_bwd(dy.contiguous(), x, bias, ctx.relu, res, gamma, stats,
                                                                want_dres=has_res and ctx.needs_input_grad[2])
        return (dx if ctx.needs_input_grad[0] else None,
                dbias if has_bias and ctx.needs_input_grad[1] else None,
                dres if has_res and ctx.needs_input_grad[2] else None,
                dgamma if has_gamma and ctx.needs_input_grad[3] else None,
                dbeta if has_beta and ctx.needs_input_grad[4] else None,
                None, None)


def bias_act_norm(x: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
                  residual: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                  beta: Optional[torch.Tensor] = None, eps: float = 1e-5) -> torch.Tensor:
    """Differentiable LayerNorm(residual + relu(x + bias)), every stage optional; x [rows, d] fp32 CUDA.
    The fused backward needs d in {128, 256, 512, 1024}; callers check `fused_dense_ok` first."""
    if not torch.is_grad_enabled():
        return _lib.bias_act_norm(x.contiguous(), bias, relu, None if residual is None else residual.contiguous(),
                                  gamma, beta, eps)
    return _BiasActNorm.apply(x, bias, residual, gamma, beta, relu, eps)


FUSED_DENSE_MIN_ROWS = 8192      # below this a forward is launch-bound and the ATen chain has less host overhead


def fused_dense_ok(x: torch.Tensor, width: int) -> bool:
    """Whether [rows, width] fp32 rows should take the fused dense glue under the current autograd mode."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.shape[0] < FUSED_DENSE_MIN_ROWS:
        return False
    return (not torch.is_grad_enabled()) or width in _lib.BIAS_ACT_NORM_BWD_WIDTHS


# ----------------------------------------------------------------------------------------------------------------------
# Training-path dense glue: rowop (fused forward AND backward, bf16 / fp32 rows, dropout) + bias-free Linear on the
# tensor cores.  An MLP in bf16 mode is   rowop(LN) -> linear_nb -> rowop(bias, ReLU, LN, dropout) -> linear_nb ->
# rowop(bias, ReLU, dropout)   forward and the mirrored chain backward, with bf16 activations in between.
# ----------------------------------------------------------------------------------------------------------------------
def _new_seed() -> int:
    """Dropout seed drawn from torch's CPU generator: follows torch.manual_seed, no device sync."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


class _RowOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, residual, gamma, beta, relu, eps, relu_out, drop_p, out_dtype):
        x = x.contiguous()
        res = None if residual is None else residual.contiguous()
        seed = _new_seed() if drop_p > 0 else 0
        out, stats = _lib.rowop_fwd(x, bias, relu, res, gamma, beta, eps, relu_out, drop_p, seed, out_dtype,
                                    want_stats=True)
        ctx.cfg = (relu, relu_out, drop_p, seed)
        ctx.has = (bias is not None, residual is not None, gamma is not None, beta is not None)
        ctx.save_for_backward(x, bias, res, gamma, beta, stats)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, bias, res, gamma, beta, stats = ctx.saved_tensors
        relu, relu_out, drop_p, seed = ctx.cfg
        has_bias, has_res, has_gamma, has_beta = ctx.has
        dx, dres, dgamma, dbeta, dbias = _lib.rowop_bwd(dy.contiguous(), x, bias, relu, res, gamma, beta, stats, relu_out,
                                                        drop_p, seed, want_dres=has_res and ctx.needs_input_grad[2])
        if dx.dtype != x.dtype:
            dx = dx.to(x.dtype)
        if dres is not None and dres.dtype != res.dtype:
            dres = dres.to(res.dtype)
        return (dx if ctx.needs_input_grad[0] else None,
                dbias if has_bias and ctx.needs_input_grad[1] else None,
                dres if has_res and ctx.needs_input_grad[2] else None,
                dgamma if has_gamma and ctx.needs_input_grad[3] else None,
                dbeta if has_beta and ctx.needs_input_grad[4] else None,
                None, None, None, None, None)


def rowop_ok(x: torch.Tensor) -> bool:
    """[rows, d] rows the rowop kernels take: CUDA, fp32 / bf16, d in {64,128,256,512,1024}."""
    return (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16)
            and x.shape[1] in _lib.ROWOP_WIDTHS)


def rowop(x: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
          residual: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
          beta: Optional[torch.Tensor] = None, eps: float = 1e-5, relu_out: bool = False, drop_p: float = 0.0,
          out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Differentiable  dropout_p(relu_out(LayerNorm(residual + relu(x + bias))))  in one pass (reference
    src/layers.py:571-579 between two Linears; :153-157 around PMA's rFF).  Widths the kernels do not take are composed
    from ATen ops ON THE SAME DEVICE (odd widths such as a dataset's raw feature count or the class count)."""
    if residual is not None and residual.dtype != x.dtype:
        residual = residual.to(x.dtype)
    if not rowop_ok(x):
        y = x.float()
        if bias is not None:
            y = y + bias
        if relu:
            y = torch.relu(y)
        if residual is not None:
            y = y + residual.float()
        if gamma is not None:
            y = torch.nn.functional.layer_norm(y, (y.shape[-1],), gamma, beta, eps)
        if relu_out:
            y = torch.relu(y)
        if drop_p > 0:
            y = torch.nn.functional.dropout(y, drop_p, True)
        return y.to(out_dtype or x.dtype)
    if not torch.is_grad_enabled():
        seed = _new_seed() if drop_p > 0 else 0
        return _lib.rowop_fwd(x.contiguous(), bias, relu, None if residual is None else residual.contiguous(), gamma, beta,
                              eps, relu_out, drop_p, seed, out_dtype)
    return _RowOp.apply(x, bias, residual, gamma, beta, relu, eps, relu_out, drop_p, out_dtype)


TC_LINEAR = True     # square 64 / 128 Linears on the hand-written tcgen05 kernels (False: cuBLAS, for A/B runs)
TC_LINEAR_PARTS = {'fwd', 'dgrad', 'wgrad'}     # A/B switch per kernel (debugging aid)


def tc_linear_ok(x: torch.Tensor, w: torch.Tensor) -> bool:
    """x @ w.T on allset_linear_fwd / its gradients on allset_linear_fwd(transposed) + allset_linear_wgrad."""
    return TC_LINEAR and x.shape[0] >= FUSED_DENSE_MIN_ROWS and _lib.linear_ok(x, w)


class _LinearNB(torch.autograd.Function):
    """y = x @ W^T without bias, W kept as the fp32 master parameter, dW accumulated and returned in fp32.

    Square widths 64 / 128 (every Linear between the aggregation steps) run on the hand-written tcgen05 kernels: bf16 rows
    with bf16 operands, fp32 rows in split precision (three bf16 terms per operand and six products for the forward and
    the input gradient, two terms for the weight gradient: as close to fp64 as an fp32 SGEMM) -- forward
    (allset_linear_fwd), input gradient (the same kernel with the weight read transposed) and weight gradient
    (allset_linear_wgrad, MN-major operands straight from the row-major activations).  Other shapes (the first layer from
    a dataset's raw feature count, the classifier's class count, the skinny score GEMM) go to cuBLAS.  `out_fp32` keeps
    the fp32 accumulator as the result of a bf16 GEMM (attention scores)."""

    @staticmethod
    def forward(ctx, x, w, out_fp32):
        ctx.w_dtype = w.dtype
        ctx.tc = tc_linear_ok(x, w)
        if ctx.tc:
            ctx.save_for_backward(x, w)
            if 'fwd' in TC_LINEAR_PARTS:
                return _lib.linear_fwd(x, w, out_dtype=torch.float32 if out_fp32 else x.dtype)
        else:
            ctx.save_for_backward(x, w if w.dtype == x.dtype else w.to(x.dtype))
        wc = w if w.dtype == x.dtype else w.to(x.dtype)
        if out_fp32 and x.dtype != torch.float32:
            return torch.mm(x, wc.t(), out_dtype=torch.float32)
        return torch.mm(x, wc.t())

    @staticmethod
    def backward(ctx, dy):
        x, wc = ctx.saved_tensors
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dy = dy.contiguous()
        tc = ctx.tc and dy.data_ptr() % 32 == 0
        dx = dw = None
        if tc and 'dgrad' in TC_LINEAR_PARTS:
            dx = _lib.linear_fwd(dy, wc, transposed=True) if ctx.needs_input_grad[0] else None
        elif ctx.needs_input_grad[0]:
            dx = torch.mm(dy, wc if wc.dtype == x.dtype else wc.to(x.dtype))
        if tc and 'wgrad' in TC_LINEAR_PARTS:
            dw = _lib.linear_wgrad(dy, x) if ctx.needs_input_grad[1] else None
        elif ctx.needs_input_grad[1]:
            if dy.dtype == torch.float32:
                dw = torch.mm(dy.t(), x)
            else:
                dw = torch.mm(dy.t(), x, out_dtype=torch.float32)
            if dw.dtype != ctx.w_dtype:
                dw = dw.to(ctx.w_dtype)
        return dx, dw, None


def linear_nb(x: torch.Tensor, w: torch.Tensor, out_fp32: bool = False) -> torch.Tensor:
    """x [rows, in] (bf16 | fp32) @ w[out, in]^T -> [rows, out] in x.dtype (fp32 with `out_fp32`); the bias is left to
    the rowop that follows."""
    if not torch.is_grad_enabled():
        if tc_linear_ok(x, w):
            return _lib.linear_fwd(x, w, out_dtype=torch.float32 if out_fp32 else x.dtype)
        wc = w if w.dtype == x.dtype else w.to(x.dtype)
        if out_fp32 and x.dtype != torch.float32:
            return torch.mm(x, wc.t(), out_dtype=torch.float32)
        return torch.mm(x, wc.t())
    return _LinearNB.apply(x, w, out_fp32)


def linear_fused(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], ln=None, relu: bool = False,
                 compute_dtype: torch.dtype = torch.float32, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """No-grad  [relu](LN?(x) w^T + b)  as ONE tcgen05 launch (LayerNorm prologue, bias and ReLU inside the kernel); the
    caller checks `tc_linear_ok(x, w)`.  fp32 compute: split precision, f32 rows in and out; bf16 compute: bf16 operands,
    any row dtypes.  An eval-mode MLP is one launch per Linear instead of a GEMM and a glue pass each."""
    if compute_dtype == torch.float32:
        return _lib.linear_fwd(x if x.dtype == torch.float32 else x.float(), w, b, ln=ln, relu=relu, split=True)
    return _lib.linear_fwd(x, w, b, ln=ln, relu=relu, split=False, out_dtype=out_dtype or torch.bfloat16)


def linear_bias_act(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], relu: bool = False,
                    out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """[relu](x w^T + b) in `out_dtype`: one fused tcgen05 launch without autograd, linear_nb + rowop (each with its fused
    backward) under autograd or for shapes the kernel does not take."""
    od = out_dtype or x.dtype
    if not torch.is_grad_enabled() and tc_linear_ok(x, w) and (x.dtype == torch.bfloat16 or od == torch.float32):
        return _lib.linear_fwd(x, w, b, relu=relu, out_dtype=od)
    return rowop(linear_nb(x, w), b, relu=relu, out_dtype=od)
