"""ctypes binding of liballset_b200.so -- the C ABI declared in include/allset_b200.h.

This is the ONLY way the Python host reaches the device code.  There is no CPU or PyTorch fallback:
if the shared library is missing, `lib()` raises with the build command (`python -m allset_b200.build`).
Every wrapper takes torch CUDA tensors, checks dtype/contiguity/device, and enqueues on
`torch.cuda.current_stream()`; buffers (outputs and workspaces) are allocated by the caller-side
PyTorch caching allocator, never by the library.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ALLSET_B200_LIB') or os.path.join(_HERE, 'liballset_b200.so')   # override: kernel tuning builds
ABI_VERSION = 3

F32, BF16 = 0, 1
SUM, MEAN = 0, 1
EUNSUPPORTED = -5

_c = ctypes
_p = _c.c_void_p
_i32, _i64, _f32, _sz = _c.c_int32, _c.c_int64, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/allset_b200.h one to one
SIGNATURES = {
    'allset_version': (_c.c_int, []),
    'allset_last_error': (_c.c_char_p, []),
    'allset_stream_eligible': (_c.c_int, [_c.c_int, _i32, _i64]),
    'allset_pma_stream_eligible': (_c.c_int, [_c.c_int, _i32, _i32, _i64]),
    'allset_stream_workspace_bytes': (_sz, [_i32]),
    'allset_csr_workspace_bytes': (_sz, [_i64, _i64]),
    'allset_csr_from_coo': (_c.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    'allset_long_segments_workspace_bytes': (_sz, [_i64]),
    'allset_long_segments': (_c.c_int, [_p, _i64, _i32, _p, _p, _p, _sz, _p]),
    'allset_segreduce_fwd': (_c.c_int, [_p, _c.c_int, _i64, _i32, _p, _p, _p, _p, _i64, _c.c_int,
                                        _p, _i32, _i32, _p, _p, _sz, _p]),
    'allset_segreduce_fwd_bcast': (_c.c_int, [_p, _c.c_int, _i64, _i32, _p, _p, _p, _p, _i64, _c.c_int,
                                              _p, _c.POINTER(_c.c_void_p), _i32, _p, _p, _sz, _p]),
    'allset_push_rows': (_c.c_int, [_p, _i64, _i64, _c.POINTER(_c.c_void_p), _i32, _p, _p]),
    'allset_bias_act_norm': (_c.c_int, [_p, _p, _c.c_int, _p, _p, _p, _f32, _i64, _i32, _p, _p, _p]),
    'allset_bias_act_norm_bwd_blocks': (_i32, [_i64]),
    'allset_bias_act_norm_bwd': (_c.c_int, [_p, _p, _p, _c.c_int, _p, _p, _p, _i64, _i32, _p, _p, _p, _p]),
    'allset_rowop_supported': (_c.c_int, [_i32]),
    'allset_rowop_fwd': (_c.c_int, [_p, _c.c_int, _p, _c.c_int, _p, _p, _p, _f32, _c.c_int, _f32, _c.c_uint64, _i64, _i32,
                                    _p, _c.c_int, _p, _p]),
    'allset_rowop_bwd': (_c.c_int, [_p, _c.c_int, _p, _c.c_int, _p, _c.c_int, _p, _p, _p, _p, _c.c_int, _f32, _c.c_uint64,
                                    _i64, _i32, _p, _p, _p, _p]),
    'allset_mlp2_fwd': (_c.c_int, [_p, _c.c_int, _p, _p, _f32, _p, _p, _p, _p, _f32, _p, _p, _c.c_int, _i64, _i32,
                                   _p, _c.c_int, _i64, _p, _p]),
    'allset_pma_fwd_strided': (_c.c_int, [_p, _i64, _p, _i64, _p, _c.c_int, _i32, _i32, _f32, _p, _p, _i64, _p, _p, _p, _sz,
                                          _p]),
    'allset_pma_tail_fwd': (_c.c_int, [_p, _c.c_int, _p, _p, _f32, _p, _p, _p, _p, _p, _p, _f32, _c.c_int, _i64, _i32,
                                       _p, _c.c_int, _p, _p]),
    'allset_linear_score_fwd': (_c.c_int, [_p, _c.c_int, _p, _p, _p, _p, _i32, _i64, _i32, _p, _c.c_int, _i64, _p, _p, _p]),
    'allset_linear_fwd': (_c.c_int, [_p, _c.c_int, _p, _p, _f32, _p, _c.c_int, _p, _c.c_int, _c.c_int, _i64, _i32, _p, _c.c_int,
                                     _p, _p]),
    'allset_linear_wgrad_partials': (_c.c_int, [_i64]),
    'allset_linear_wgrad': (_c.c_int, [_p, _p, _c.c_int, _c.c_int, _i64, _i32, _p, _p, _i64, _p, _p]),
    'allset_segreduce_bwd_w': (_c.c_int, [_p, _p, _c.c_int, _i32, _p, _p, _p, _i64, _p, _p]),
    'allset_pma_fwd': (_c.c_int, [_p, _p, _p, _c.c_int, _i32, _i32, _f32, _p, _p, _i64,
                                  _p, _i32, _i32, _p, _p, _p, _sz, _p]),
    'allset_pma_fwd_bcast': (_c.c_int, [_p, _p, _p, _c.c_int, _i32, _i32, _f32, _p, _p, _i64,
                                        _p, _p, _c.POINTER(_c.c_void_p), _i32, _p, _p, _sz, _p]),
    'allset_pma_alpha': (_c.c_int, [_p, _p, _i32, _f32, _p, _p, _i64, _p, _p]),
    'allset_rowdot_heads': (_c.c_int, [_p, _p, _p, _c.c_int, _i64, _i32, _i32, _p, _p]),
    'allset_pma_bwd': (_c.c_int, [_p, _p, _p, _p, _p, _c.c_int, _i32, _i32, _f32, _p, _p, _i64,
                                  _p, _i32, _i32, _p, _p, _p]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            'allset_b200: %s not found. The aggregation path has no CPU fallback; build the sm_100a library '
            'with `python -m allset_b200.build` (needs nvcc).' % LIB_PATH)
    h = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(h, name)            # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    if h.allset_version() != ABI_VERSION:
        raise RuntimeError('allset_b200: ABI version %d != expected %d; rebuild with `python -m allset_b200.build`'
                           % (h.allset_version(), ABI_VERSION))
    _lib = h
    return h


def _check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError('%s failed (%d): %s' % (what, code, lib().allset_last_error().decode()))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError('allset_b200 kernels store features as float32 or bfloat16, got %s' % t.dtype)


def _need(t: Optional[torch.Tensor], name: str, dtype=None, optional=False) -> None:
    if t is None:
        if optional:
            return
        raise ValueError('%s is required' % name)
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor: allset_b200 has no CPU path (got device %s)' % (name, t.device))
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError('%s must be %s, got %s' % (name, dtype, t.dtype))


# ----------------------------------------------------------------------------------------------------------
# incidence container
# ----------------------------------------------------------------------------------------------------------
def csr_from_coo(tgt: torch.Tensor, src: torch.Tensor, n_tgt: int):
    """Stable sort of a COO incidence list by target -> (rowptr[n_tgt+1], col[nnz], perm[nnz]) int32."""
    _need(tgt, 'tgt', torch.int64)
    _need(src, 'src', torch.int64)
    nnz = tgt.numel()
    dev = tgt.device
    with torch.cuda.device(dev):
        rowptr = torch.empty(n_tgt + 1, dtype=torch.int32, device=dev)
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        perm = torch.empty(nnz, dtype=torch.int32, device=dev)
        ws_bytes = lib().allset_csr_workspace_bytes(nnz, n_tgt)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _check(lib().allset_csr_from_coo(_ptr(tgt), _ptr(src), nnz, n_tgt, _ptr(rowptr), _ptr(col), _ptr(perm),
                                         _ptr(ws), ws_bytes, _stream()), 'allset_csr_from_coo')
    return rowptr, col, perm


def long_segments(rowptr: torch.Tensor, n_tgt: int, threshold: int) -> Optional[torch.Tensor]:
    """Row ids of segments longer than `threshold` (int32, ascending) or None.  Synchronises once (graph build)."""
    _need(rowptr, 'rowptr', torch.int32)
    dev = rowptr.device
    if n_tgt == 0:
        return None
    with torch.cuda.device(dev):
        ids = torch.empty(n_tgt, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = lib().allset_long_segments_workspace_bytes(n_tgt)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _check(lib().allset_long_segments(_ptr(rowptr), n_tgt, threshold, _ptr(ids), _ptr(cnt), _ptr(ws), ws_bytes,
                                          _stream()), 'allset_long_segments')
        n = int(cnt.item())
    return ids[:n].clone() if n > 0 else None


# ----------------------------------------------------------------------------------------------------------
# AllDeepSets
# ----------------------------------------------------------------------------------------------------------
# Workspace of the stream kernels (allset_stream_workspace_bytes): partial results + ready flags of the pieces of long
# segments that are cut at warp-chunk boundaries.  Zero-initialised once, left zeroed by every launch; one per
# (device, stream) because concurrent launches must not share it.
_stream_ws = {}


def stream_workspace(device: torch.device, d: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    need = lib().allset_stream_workspace_bytes(int(d))
    ws = _stream_ws.get(key)
    if ws is None or ws.numel() < need:
        with torch.cuda.device(device):
            ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _stream_ws[key] = ws
    return ws


def stream_eligible(dtype: torch.dtype, d: int, n_tgt: int, heads: int = 0) -> bool:
    """Whether [*, d] rows of `dtype` reduced into n_tgt segments take the stream kernels (heads > 0: the PMA one).  With
    their workspace those kernels handle segments of any length, so no long-segment bucketing is needed then."""
    code = F32 if dtype == torch.float32 else BF16 if dtype == torch.bfloat16 else None
    if code is None or n_tgt <= 0 or d <= 0:
        return False
    if heads > 0:
        return d % heads == 0 and bool(lib().allset_pma_stream_eligible(code, heads, d // heads, n_tgt))
    return bool(lib().allset_stream_eligible(code, d, n_tgt))


def fused_exchange_eligible(dtype: torch.dtype, d: int, n_tgt: int, heads: int = 0) -> bool:
    """Whether a rank whose range has `n_tgt` target rows takes the fused-exchange stream kernel for [*, d] rows of
    `dtype`.  A pure function of its arguments: every rank evaluates it for EVERY rank's range and fuses only if all do
    (allset_b200.sharding.ShardedIncidence.fused_ok)."""
    return stream_eligible(dtype, d, n_tgt, heads)


def segreduce_fwd(x: torch.Tensor, rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int, mean: bool,
                  w: Optional[torch.Tensor] = None, src_scale: Optional[torch.Tensor] = None,
                  long_ids: Optional[torch.Tensor] = None, long_threshold: int = 0,
                  out: Optional[torch.Tensor] = None, max_segment_len: int = 0, allow_stream: bool = True
                  ) -> torch.Tensor:
    _need(x, 'x')
    _need(rowptr, 'rowptr', torch.int32)
    _need(col, 'col', torch.int32)
    _need(w, 'w', torch.float32, optional=True)
    _need(src_scale, 'src_scale', torch.float32, optional=True)
    if x.dim() != 2:
        raise ValueError('x must be [n_src, d]')
    n_src, d = x.shape
    if rowptr.numel() < n_tgt + 1:
        raise ValueError('rowptr has %d entries, need %d' % (rowptr.numel(), n_tgt + 1))
    if w is not None and w.numel() != col.numel():
        raise ValueError('w must have one entry per incidence')
    if src_scale is not None and src_scale.numel() != n_src:
        raise ValueError('src_scale must have one entry per source row')
    if out is None:
        out = torch.empty((n_tgt, d), dtype=x.dtype, device=x.device)
    else:
        _need(out, 'out', x.dtype)
        if tuple(out.shape) != (n_tgt, d):
            raise ValueError('out must be [n_tgt, d]')
    if d == 0 or n_tgt == 0:
        return out
    ws = None
    if allow_stream and stream_eligible(x.dtype, d, n_tgt):
        ws, long_ids = stream_workspace(x.device, d), None          # the stream kernel cuts long segments itself

    n_long = 0 if long_ids is None else long_ids.numel()
    with torch.cuda.device(x.device):
        _check(lib().allset_segreduce_fwd(_ptr(x), _dtype_code(x), n_src, d, _ptr(rowptr), _ptr(col), _ptr(w),
                                          _ptr(src_scale), n_tgt, MEAN if mean else SUM, _ptr(long_ids), n_long,
                                          long_threshold, _ptr(out), _ptr(ws), 0 if ws is None else ws.numel(), _stream()),
               'allset_segreduce_fwd')
    return out


def bias_act_norm(x: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
                  residual: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                  beta: Optional[torch.Tensor] = None, eps: float = 1e-5, want_stats: bool = False):
    """LayerNorm(residual + relu(x + bias)), every stage optional, one pass over [rows, d] fp32 rows.
    want_stats: also return the per-row (mean, rstd) [rows, 2] the backward pass needs."""
    _need(x, 'x', torch.float32)
    _need(bias, 'bias', torch.float32, optional=True)
    _need(residual, 'residual', torch.float32, optional=True)
    _need(gamma, 'gamma', torch.float32, optional=True)
    _need(beta, 'beta', torch.float32, optional=True)
    if x.dim() != 2:
        raise ValueError('x must be [rows, d]')
    rows, d = x.shape
    if (bias is not None and bias.numel() != d) or (gamma is not None and gamma.numel() != d) or \
            (beta is not None and beta.numel() != d) or (residual is not None and residual.shape != x.shape):
        raise ValueError('bias_act_norm: shape mismatch')
    out = torch.empty_like(x)
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if (want_stats and gamma is not None) else None
    if rows > 0:
        with torch.cuda.device(x.device):
            _check(lib().allset_bias_act_norm(_ptr(x), _ptr(bias), 1 if relu else 0, _ptr(residual), _ptr(gamma),
                                              _ptr(beta), float(eps), rows, d, _ptr(out), _ptr(stats), _stream()),
                   'allset_bias_act_norm')
    return (out, stats) if want_stats else out


ROWOP_WIDTHS = (64, 128, 256, 512, 1024)


def rowop_fwd(x: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
              residual: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
              beta: Optional[torch.Tensor] = None, eps: float = 1e-5, relu_out: bool = False, drop_p: float = 0.0,
              seed: int = 0, out_dtype: Optional[torch.dtype] = None, want_stats: bool = False,
              out: Optional[torch.Tensor] = None):
    """out = dropout_p(relu_out(LayerNorm(residual + relu(x + bias)))), every stage optional, ONE pass; x / residual
    fp32 or bf16 (same dtype), out fp32 or bf16, parameters and statistics fp32 (allset_rowop_fwd)."""
    _need(x, 'x')
    _need(bias, 'bias', torch.float32, optional=True)
    _need(residual, 'residual', x.dtype, optional=True)
    _need(gamma, 'gamma', torch.float32, optional=True)
    _need(beta, 'beta', torch.float32, optional=True)
    if x.dim() != 2:
        raise ValueError('x must be [rows, d]')
    rows, d = x.shape
    if (bias is not None and bias.numel() != d) or (gamma is not None and gamma.numel() != d) or \
            (beta is not None and beta.numel() != d) or (residual is not None and residual.shape != x.shape):
        raise ValueError('rowop: shape mismatch')
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty((rows, d), dtype=out_dtype, device=x.device)
    else:
        _need(out, 'out', out_dtype)
        if tuple(out.shape) != (rows, d):
            raise ValueError('out must be [rows, d]')
    stats = torch.empty((rows, 2), dtype=torch.float32, device=x.device) if (want_stats and gamma is not None) else None
    if rows > 0:
        with torch.cuda.device(x.device):
            _check(lib().allset_rowop_fwd(_ptr(x), _dtype_code(x), _ptr(bias), 1 if relu else 0, _ptr(residual),
                                          _ptr(gamma), _ptr(beta), float(eps), 1 if relu_out else 0, float(drop_p),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, rows, d, _ptr(out), _dtype_code(out),
                                          _ptr(stats), _stream()), 'allset_rowop_fwd')
    return (out, stats) if want_stats else out


def rowop_bwd(dy: torch.Tensor, x: torch.Tensor, bias: Optional[torch.Tensor], relu: bool,
              residual: Optional[torch.Tensor], gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
              stats: Optional[torch.Tensor], relu_out: bool, drop_p: float, seed: int, want_dres: bool):
    """-> (dx, dres | None, dgamma, dbeta, dbias): dx / dres in dy's dtype; the column sums (fp32) come from per-CTA
    partials summed here (deterministic)."""
    _need(dy, 'dy')
    _need(x, 'x')
    rows, d = x.shape
    if tuple(dy.shape) != (rows, d):
        raise ValueError('dy must match x')
    dx = torch.empty((rows, d), dtype=dy.dtype, device=x.device)
    dres = torch.empty_like(dx) if want_dres else None
    if rows == 0:
        z = torch.zeros(d, device=x.device)
        return dx, dres, z, z.clone(), z.clone()
    blocks = lib().allset_bias_act_norm_bwd_blocks(rows)
    partial = torch.empty((blocks, 3, d), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().allset_rowop_bwd(_ptr(dy), _dtype_code(dy), _ptr(x), _dtype_code(x), _ptr(bias), 1 if relu else 0,
                                      _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(stats), 1 if relu_out else 0,
                                      float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, rows, d, _ptr(dx), _ptr(dres),
                                      _ptr(partial), _stream()), 'allset_rowop_bwd')
    sums = partial.sum(dim=0)
    return dx, dres, sums[0], sums[1], sums[2]


BIAS_ACT_NORM_BWD_WIDTHS = (128, 256, 512, 1024)


def bias_act_norm_bwd(dy: torch.Tensor, x: torch.Tensor, bias: Optional[torch.Tensor], relu: bool,
                      residual: Optional[torch.Tensor], gamma: Optional[torch.Tensor], stats: Optional[torch.Tensor],
                      want_dres: bool):
    """-> (dx, dres | None, dgamma, dbeta, dbias) ; the three column sums come from per-CTA partials summed here."""
    _need(dy, 'dy', torch.float32)
    _need(x, 'x', torch.float32)
    rows, d = x.shape
    dx = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    blocks = lib().allset_bias_act_norm_bwd_blocks(rows)
    partial = torch.empty((blocks, 3, d), dtype=torch.float32, device=x.device)
    if rows == 0:
        z = torch.zeros(d, device=x.device)
        return dx, dres, z, z.clone(), z.clone()
    with torch.cuda.device(x.device):
        _check(lib().allset_bias_act_norm_bwd(_ptr(dy), _ptr(x), _ptr(bias), 1 if relu else 0, _ptr(residual), _ptr(gamma),
                                              _ptr(stats), rows, d, _ptr(dx), _ptr(dres), _ptr(partial), _stream()),
               'allset_bias_act_norm_bwd')
    sums = partial.sum(dim=0)
    return dx, dres, sums[0], sums[1], sums[2]


MLP2_WIDTHS = (64, 128)


def mlp2_fwd(x: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: Optional[torch.Tensor],
             b2: Optional[torch.Tensor], ln0=None, ln1=None, relu_out: bool = False,
             out_dtype: Optional[torch.dtype] = None, status: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = [relu]( LN1?( relu( LN0?(x) W1^T + b1 ) ) W2^T + b2 ) in one tcgen05 kernel (bf16 operands, fp32
    accumulate).  x [rows, d] f32|bf16, d in MLP2_WIDTHS; w1, w2 [d, d] f32 ([out, in]); ln0 / ln1 = None or
    (gamma, beta | None, eps); out_dtype f32 (default) | bf16.  `status`: optional int32[1] diagnostic word.
    w2 is None: ONE Linear, out = [relu](LN0?(x) W1^T + b1) (b2 / ln1 must be None)."""
    _need(x, 'x')
    xd = _dtype_code(x)
    if x.dim() != 2:
        raise ValueError('x must be [rows, d]')
    rows, d = x.shape
    if w2 is None and (b2 is not None or ln1 is not None):
        raise ValueError('mlp2_fwd: w2 is None (single Linear) excludes b2 / ln1')
    for name, t, shape in (('w1', w1, (d, d)), ('w2', w2, (d, d))):
        if t is None and name == 'w2':
            continue
        _need(t, name, torch.float32)
        if tuple(t.shape) != shape:
            raise ValueError('mlp2_fwd: %s must be %s, got %s' % (name, shape, tuple(t.shape)))
    for name, t in (('b1', b1), ('b2', b2)):
        _need(t, name, torch.float32, optional=True)
        if t is not None and t.numel() != d:
            raise ValueError('mlp2_fwd: %s must have %d entries' % (name, d))
    g = [None, None]
    b = [None, None]
    eps = [1e-5, 1e-5]
    for i, ln in enumerate((ln0, ln1)):
        if ln is None:
            continue
        g[i], b[i], eps[i] = ln
        _need(g[i], 'ln%d gamma' % i, torch.float32)
        _need(b[i], 'ln%d beta' % i, torch.float32, optional=True)
        if g[i].numel() != d or (b[i] is not None and b[i].numel() != d):
            raise ValueError('mlp2_fwd: LayerNorm %d parameters must have %d entries' % (i, d))
    if out is None:
        out_dtype = torch.float32 if out_dtype is None else out_dtype
        out = torch.empty((rows, d), dtype=out_dtype, device=x.device)
        pitch = 0
    else:
        # a [rows, d] view with unit column stride and a row pitch (e.g. the value part of packed PMA records)
        if not out.is_cuda or tuple(out.shape) != (rows, d) or out.stride(1) != 1:
            raise ValueError('mlp2_fwd: out must be a CUDA [rows, d] view with unit column stride')
        pitch = out.stride(0) * out.element_size()
    od = _dtype_code(out)
    _need(status, 'status', torch.int32, optional=True)
    with torch.cuda.device(x.device):
        _check(lib().allset_mlp2_fwd(_ptr(x), xd, _ptr(g[0]), _ptr(b[0]), float(eps[0]), _ptr(w1), _ptr(b1),
                                     _ptr(g[1]), _ptr(b[1]), float(eps[1]), _ptr(w2), _ptr(b2), 1 if relu_out else 0,
                                     rows, d, _ptr(out), od, pitch, _ptr(status), _stream()), 'allset_mlp2_fwd')
    return out


PREC_BF16, PREC_SPLIT = 0, 1
LINEAR_WIDTHS = (64, 128)


def linear_ok(x: torch.Tensor, w: torch.Tensor) -> bool:
    """Whether `x @ w.T` (or its gradients) can run on the hand-written tcgen05 Linear kernels: square width 64 / 128,
    dense CUDA rows in f32 (split precision) or bf16, f32 master weights."""
    return (x.is_cuda and x.dim() == 2 and x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
            and w.dim() == 2 and w.dtype == torch.float32 and w.is_contiguous()
            and x.shape[1] in LINEAR_WIDTHS and tuple(w.shape) == (x.shape[1], x.shape[1])
            and x.data_ptr() % 32 == 0 and w.data_ptr() % 16 == 0)


def linear_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, ln=None, relu: bool = False,
               transposed: bool = False, out_dtype: Optional[torch.dtype] = None,
               status: Optional[torch.Tensor] = None, split: Optional[bool] = None) -> torch.Tensor:
    """out = [relu](LN?(x) op(w)^T + b) on tcgen05; op(w) = w ([out, in]) or w^T with `transposed` (dx = dy w).
    f32 rows run in split precision (three bf16 terms per operand, six products: fp32 accuracy) and return f32 unless
    `split=False`; bf16 rows (and f32 rows with `split=False`) run with bf16 operands and return `out_dtype` (default: the
    input dtype)."""
    _need(x, 'x')
    xd = _dtype_code(x)
    if x.dim() != 2:
        raise ValueError('x must be [rows, d]')
    rows, d = x.shape
    _need(w, 'w', torch.float32)
    if tuple(w.shape) != (d, d):
        raise ValueError('linear_fwd: w must be %s, got %s' % ((d, d), tuple(w.shape)))
    _need(b, 'b', torch.float32, optional=True)
    g = bt = None
    eps = 1e-5
    if ln is not None:
        g, bt, eps = ln
        _need(g, 'ln gamma', torch.float32)
        _need(bt, 'ln beta', torch.float32, optional=True)
    if split is None:
        split = x.dtype == torch.float32          # f32 rows: the reference's precision class unless told otherwise
    if split:
        if x.dtype != torch.float32:
            raise ValueError('linear_fwd: split precision takes f32 rows')
        if out_dtype not in (None, torch.float32):
            raise ValueError('linear_fwd: f32 rows (split precision) return f32')
        out_dtype = torch.float32
    out = torch.empty((rows, d), dtype=out_dtype or x.dtype, device=x.device)
    _need(status, 'status', torch.int32, optional=True)
    with torch.cuda.device(x.device):
        _check(lib().allset_linear_fwd(_ptr(x), xd, _ptr(g), _ptr(bt), float(eps), _ptr(w), 1 if transposed else 0, _ptr(b),
                                       1 if relu else 0, PREC_SPLIT if split else PREC_BF16, rows, d, _ptr(out),
                                       _dtype_code(out), _ptr(status), _stream()), 'allset_linear_fwd')
    return out


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw[n, k] = sum_r dy[r, n] x[r, k]  ([d, d] f32) on tcgen05; dy, x [rows, d] of one dtype (f32: split precision)."""
    _need(dy, 'dy')
    _need(x, 'x')
    if dy.dtype != x.dtype or dy.shape != x.shape or x.dim() != 2:
        raise ValueError('linear_wgrad: dy and x must be [rows, d] of one dtype')
    rows, d = x.shape
    dw = torch.empty((d, d), dtype=torch.float32, device=x.device)
    _need(status, 'status', torch.int32, optional=True)
    with torch.cuda.device(x.device):
        n_part = int(lib().allset_linear_wgrad_partials(rows))
        ws = torch.empty((max(n_part, 1), d, d), dtype=torch.float32, device=x.device)
        _check(lib().allset_linear_wgrad(_ptr(dy), _ptr(x), _dtype_code(x), PREC_SPLIT if x.dtype == torch.float32 else PREC_BF16,
                                         rows, d, _ptr(dw), _ptr(ws), ws.numel(), _ptr(status), _stream()),
               'allset_linear_wgrad')
    return dw


def pma_tail_fwd(x: torch.Tensor, ln0, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
                 b2: Optional[torch.Tensor], ln1, relu_final: bool = False,
                 out_dtype: Optional[torch.dtype] = None, status: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = [relu]( LN1( y + relu( relu(y W1^T + b1) W2^T + b2 ) ) ), y = LN0(x): PMA's ln0 / rFF / residual / ln1 in one
    tcgen05 kernel.  x [rows, d] f32|bf16, d in MLP2_WIDTHS; ln0, ln1 = (gamma, beta | None, eps)."""
    _need(x, 'x')
    xd = _dtype_code(x)
    if x.dim() != 2:
        raise ValueError('x must be [rows, d]')
    rows, d = x.shape
    for name, t in (('w1', w1), ('w2', w2)):
        _need(t, name, torch.float32)
        if tuple(t.shape) != (d, d):
            raise ValueError('pma_tail_fwd: %s must be %s, got %s' % (name, (d, d), tuple(t.shape)))
    for name, t in (('b1', b1), ('b2', b2), ('ln0 gamma', ln0[0]), ('ln0 beta', ln0[1]), ('ln1 gamma', ln1[0]),
                    ('ln1 beta', ln1[1])):
        _need(t, name, torch.float32, optional=name not in ('ln0 gamma', 'ln1 gamma'))
        if t is not None and t.numel() != d:
            raise ValueError('pma_tail_fwd: %s must have %d entries' % (name, d))
    out = torch.empty((rows, d), dtype=torch.float32 if out_dtype is None else out_dtype, device=x.device)
    od = _dtype_code(out)
    _need(status, 'status', torch.int32, optional=True)
    with torch.cuda.device(x.device):
        _check(lib().allset_pma_tail_fwd(_ptr(x), xd, _ptr(ln0[0]), _ptr(ln0[1]), float(ln0[2]), _ptr(w1), _ptr(b1),
                                         _ptr(w2), _ptr(b2), _ptr(ln1[0]), _ptr(ln1[1]), float(ln1[2]),
                                         1 if relu_final else 0, rows, d, _ptr(out), od, _ptr(status), _stream()),
               'allset_pma_tail_fwd')
    return out


def linear_score_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], w_eff: torch.Tensor,
                     b_eff: Optional[torch.Tensor], out_dtype: Optional[torch.dtype] = None,
                     out: Optional[torch.Tensor] = None, status: Optional[torch.Tensor] = None):
    """-> (out = x W^T + b  [rows, d] (tcgen05, bf16 operands), score = x w_eff^T + b_eff  [rows, H] f32 (fp32 FMAs)) in
    one launch: PMA.lin_V and the folded lin_K score.  `out` may be a pitched [rows, d] view (see mlp2_fwd)."""
    _need(x, 'x')
    xd = _dtype_code(x)
    rows, d = x.shape
    _need(w, 'w', torch.float32)
    _need(b, 'b', torch.float32, optional=True)
    _need(w_eff, 'w_eff', torch.float32)
    _need(b_eff, 'b_eff', torch.float32, optional=True)
    H = w_eff.shape[0]
    if tuple(w.shape) != (d, d) or w_eff.dim() != 2 or w_eff.shape[1] != d or (b_eff is not None and b_eff.numel() != H):
        raise ValueError('linear_score_fwd: shape mismatch')
    if out is None:
        out = torch.empty((rows, d), dtype=torch.float32 if out_dtype is None else out_dtype, device=x.device)
        pitch = 0
    else:
        if not out.is_cuda or tuple(out.shape) != (rows, d) or out.stride(1) != 1:
            raise ValueError('linear_score_fwd: out must be a CUDA [rows, d] view with unit column stride')
        pitch = out.stride(0) * out.element_size()
    score = torch.empty((rows, H), dtype=torch.float32, device=x.device)
    _need(status, 'status', torch.int32, optional=True)
    with torch.cuda.device(x.device):
        _check(lib().allset_linear_score_fwd(_ptr(x), xd, _ptr(w), _ptr(b), _ptr(w_eff), _ptr(b_eff), H, rows, d, _ptr(out),
                                             _dtype_code(out), pitch, _ptr(score), _ptr(status), _stream()),
               'allset_linear_score_fwd')
    return out, score


class Unsupported(RuntimeError):
    """The fused-exchange variant does not handle this shape; fall back to the plain call + an all-gather."""


def _peer_array(peer_ptrs):
    arr = (_c.c_void_p * max(len(peer_ptrs), 1))()
    for i, q in enumerate(peer_ptrs):
        arr[i] = int(q)
    return arr


def segreduce_fwd_bcast(x: torch.Tensor, rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int, mean: bool,
                        out: torch.Tensor, peer_ptrs, w: Optional[torch.Tensor] = None,
                        src_scale: Optional[torch.Tensor] = None, peer_mask: Optional[torch.Tensor] = None
                        ) -> torch.Tensor:
    """segreduce_fwd whose epilogue also sends every reduced row to the same row of each peer replica.
    `out` = this rank's row range inside its own replica; `peer_ptrs` = addresses (ints, peer-mapped device pointers)
    of that same row range in the other ranks' replicas; `peer_mask` [n_tgt] uint8 (bit j = peer j needs the row) or
    None = all.  Raises Unsupported for shapes the stream kernel rejects."""
    _need(x, 'x')
    _need(rowptr, 'rowptr', torch.int32)
    _need(col, 'col', torch.int32)
    _need(out, 'out', x.dtype)
    _need(w, 'w', torch.float32, optional=True)
    _need(src_scale, 'src_scale', torch.float32, optional=True)
    _need(peer_mask, 'peer_mask', torch.uint8, optional=True)
    n_src, d = x.shape
    if tuple(out.shape) != (n_tgt, d):
        raise ValueError('out must be [n_tgt, d]')
    if peer_mask is not None and peer_mask.numel() != n_tgt:
        raise ValueError('peer_mask must have one byte per target row')
    if n_tgt == 0:
        return out
    ws = stream_workspace(x.device, d)
    with torch.cuda.device(x.device):
        code = lib().allset_segreduce_fwd_bcast(_ptr(x), _dtype_code(x), n_src, d, _ptr(rowptr), _ptr(col), _ptr(w),
                                                _ptr(src_scale), n_tgt, MEAN if mean else SUM, _ptr(out),
                                                _peer_array(peer_ptrs), len(peer_ptrs), _ptr(peer_mask), _ptr(ws),
                                                ws.numel(), _stream())
    if code == EUNSUPPORTED:
        raise Unsupported(lib().allset_last_error().decode())
    _check(code, 'allset_segreduce_fwd_bcast')
    return out


def pma_fwd_bcast(v: torch.Tensor, score: torch.Tensor, seed: torch.Tensor, H: int, C: int, slope: float,
                  rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int, out: torch.Tensor, peer_ptrs,
                  stats: Optional[torch.Tensor] = None, peer_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(v, 'v')
    _need(score, 'score', torch.float32)
    _need(seed, 'seed', torch.float32)
    _need(out, 'out', v.dtype)
    _need(stats, 'stats', torch.float32, optional=True)
    _need(peer_mask, 'peer_mask', torch.uint8, optional=True)
    if tuple(out.shape) != (n_tgt, H * C):
        raise ValueError('out must be [n_tgt, H*C]')
    if n_tgt == 0:
        return out
    ws = stream_workspace(v.device, H * C)
    with torch.cuda.device(v.device):
        code = lib().allset_pma_fwd_bcast(_ptr(v), _ptr(score), _ptr(seed), _dtype_code(v), H, C, float(slope),
                                          _ptr(rowptr), _ptr(col), n_tgt, _ptr(out), _ptr(stats),
                                          _peer_array(peer_ptrs), len(peer_ptrs), _ptr(peer_mask), _ptr(ws), ws.numel(),
                                          _stream())
    if code == EUNSUPPORTED:
        raise Unsupported(lib().allset_last_error().decode())
    _check(code, 'allset_pma_fwd_bcast')
    return out


def push_rows(rows: torch.Tensor, peer_ptrs, peer_mask: Optional[torch.Tensor] = None) -> None:
    """Send `rows` (this rank's [n, d] row range, a view into its own replica) to the same rows of the peer replicas
    (`peer_ptrs`: peer-mapped addresses of that range); `peer_mask` [n] uint8 selects the peers per row."""
    _need(rows, 'rows')
    _need(peer_mask, 'peer_mask', torch.uint8, optional=True)
    if rows.dim() != 2:
        raise ValueError('rows must be [n, d]')
    n, d = rows.shape
    if peer_mask is not None and peer_mask.numel() != n:
        raise ValueError('peer_mask must have one byte per row')
    if n == 0 or len(peer_ptrs) == 0:
        return
    with torch.cuda.device(rows.device):
        _check(lib().allset_push_rows(_ptr(rows), n, d * rows.element_size(), _peer_array(peer_ptrs), len(peer_ptrs),
                                      _ptr(peer_mask), _stream()), 'allset_push_rows')


def packed_pma_records(n_src: int, d: int, H: int, dtype: torch.dtype, device) -> tuple:
    """One [values | scores] record per source row: returns (buf uint8 [n_src, d*es + H*4], values view [n_src, d] of
    `dtype`, scores view [n_src, H] float32).  Both views have unit column stride and the record size as row pitch."""
    es = torch.empty(0, dtype=dtype).element_size()
    rowb = d * es
    buf = torch.empty((n_src, rowb + H * 4), dtype=torch.uint8, device=device)
    return buf, buf[:, :rowb].view(dtype), buf[:, rowb:].view(torch.float32)


def pma_fwd_strided(v: torch.Tensor, score: torch.Tensor, seed: torch.Tensor, H: int, C: int, slope: float,
                    rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int, want_stats: bool = False):
    """pma_fwd over strided sources (views with unit column stride and a row pitch, e.g. packed_pma_records): stream
    kernel only -- raises Unsupported where it does not apply.  -> (out [n_tgt, H*C] dense, stats | None)."""
    for name, t, cols in (('v', v, H * C), ('score', score, H)):
        if not t.is_cuda or t.dim() != 2 or t.shape[1] != cols or t.stride(1) != 1:
            raise ValueError('pma_fwd_strided: %s must be a CUDA [n_src, %d] view with unit column stride' % (name, cols))
    if score.dtype != torch.float32 or score.shape[0] != v.shape[0]:
        raise ValueError('pma_fwd_strided: score must be float32 [n_src, H]')
    _need(seed, 'seed', torch.float32)
    _need(rowptr, 'rowptr', torch.int32)
    _need(col, 'col', torch.int32)
    out = torch.empty((n_tgt, H * C), dtype=v.dtype, device=v.device)
    stats = torch.empty((n_tgt, H, 2), dtype=torch.float32, device=v.device) if want_stats else None
    if n_tgt == 0:
        return out, stats
    ws = stream_workspace(v.device, H * C)
    with torch.cuda.device(v.device):
        code = lib().allset_pma_fwd_strided(_ptr(v), v.stride(0) * v.element_size(), _ptr(score), score.stride(0) * 4,
                                            _ptr(seed), _dtype_code(v), H, C, float(slope), _ptr(rowptr), _ptr(col),
                                            n_tgt, _ptr(out), _ptr(stats), _ptr(ws), ws.numel(), _stream())
    if code == EUNSUPPORTED:
        raise Unsupported(lib().allset_last_error().decode())
    _check(code, 'allset_pma_fwd_strided')
    return out, stats


def segreduce_bwd_w(x: torch.Tensor, grad_out: torch.Tensor, rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int,
                    tgt_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(x, 'x')
    _need(grad_out, 'grad_out', x.dtype)
    _need(tgt_scale, 'tgt_scale', torch.float32, optional=True)
    gw = torch.empty(col.numel(), dtype=torch.float32, device=x.device)
    if col.numel() == 0:
        return gw
    with torch.cuda.device(x.device):
        _check(lib().allset_segreduce_bwd_w(_ptr(x), _ptr(grad_out), _dtype_code(x), x.shape[1], _ptr(rowptr), _ptr(col),
                                            _ptr(tgt_scale), n_tgt, _ptr(gw), _stream()), 'allset_segreduce_bwd_w')
    return gw


# ----------------------------------------------------------------------------------------------------------
# AllSetTransformer (PMA)
# ----------------------------------------------------------------------------------------------------------
def pma_fwd(v: torch.Tensor, score: torch.Tensor, seed: torch.Tensor, H: int, C: int, slope: float,
            rowptr: torch.Tensor, col: torch.Tensor, n_tgt: int, want_stats: bool = True,
            long_ids: Optional[torch.Tensor] = None, long_threshold: int = 0, out: Optional[torch.Tensor] = None,
            max_segment_len: int = 0, allow_stream: bool = True):
    """v [n_src, H*C], score [n_src, H] f32, seed [H*C] f32 -> (out [n_tgt, H*C], stats [n_tgt, H, 2] | None)."""
    _need(v, 'v')
    _need(score, 'score', torch.float32)
    _need(seed, 'seed', torch.float32)
    _need(rowptr, 'rowptr', torch.int32)
    _need(col, 'col', torch.int32)
    if v.dim() != 2 or v.shape[1] != H * C or score.shape != (v.shape[0], H) or seed.numel() != H * C:
        raise ValueError('pma_fwd: shape mismatch v%s score%s seed%s H=%d C=%d'
                         % (tuple(v.shape), tuple(score.shape), tuple(seed.shape), H, C))
    if out is None:
        out = torch.empty((n_tgt, H * C), dtype=v.dtype, device=v.device)
    else:
        _need(out, 'out', v.dtype)
        if tuple(out.shape) != (n_tgt, H * C):
            raise ValueError('out must be [n_tgt, H*C]')
    stats = torch.empty((n_tgt, H, 2), dtype=torch.float32, device=v.device) if want_stats else None
    if n_tgt == 0:
        return out, stats
    ws = None
    if allow_stream and stream_eligible(v.dtype, H * C, n_tgt, heads=H):
        ws, long_ids = stream_workspace(v.device, H * C), None       # the stream kernel cuts long segments itself

    n_long = 0 if long_ids is None else long_ids.numel()
    with torch.cuda.device(v.device):
        _check(lib().allset_pma_fwd(_ptr(v), _ptr(score), _ptr(seed), _dtype_code(v), H, C, float(slope), _ptr(rowptr),
                                    _ptr(col), n_tgt, _ptr(long_ids), n_long, long_threshold, _ptr(out), _ptr(stats),
                                    _ptr(ws), 0 if ws is None else ws.numel(), _stream()), 'allset_pma_fwd')
    return out, stats


def pma_alpha(score: torch.Tensor, stats: torch.Tensor, H: int, slope: float, rowptr: torch.Tensor,
              col: torch.Tensor, n_tgt: int) -> torch.Tensor:
    _need(score, 'score', torch.float32)
    _need(stats, 'stats', torch.float32)
    alpha = torch.empty((col.numel(), H), dtype=torch.float32, device=score.device)
    if col.numel() == 0:
        return alpha
    with torch.cuda.device(score.device):
        _check(lib().allset_pma_alpha(_ptr(score), _ptr(stats), H, float(slope), _ptr(rowptr), _ptr(col), n_tgt,
                                      _ptr(alpha), _stream()), 'allset_pma_alpha')
    return alpha


def rowdot_heads(a: torch.Tensor, b: torch.Tensor, sub: Optional[torch.Tensor], H: int, C: int) -> torch.Tensor:
    _need(a, 'a')
    _need(b, 'b', a.dtype)
    _need(sub, 'sub', torch.float32, optional=True)
    n = a.shape[0]
    out = torch.empty((n, H), dtype=torch.float32, device=a.device)
    if n == 0:
        return out
    with torch.cuda.device(a.device):
        _check(lib().allset_rowdot_heads(_ptr(a), _ptr(b), _ptr(sub), _dtype_code(a), n, H, C, _ptr(out), _stream()),
               'allset_rowdot_heads')
    return out


def pma_bwd(grad_out: torch.Tensor, v: torch.Tensor, score: torch.Tensor, stats: torch.Tensor, D: torch.Tensor,
            H: int, C: int, slope: float, rowptrT: torch.Tensor, colT: torch.Tensor, n_src: int,
            long_ids: Optional[torch.Tensor] = None, long_threshold: int = 0):
    _need(grad_out, 'grad_out', v.dtype)
    _need(v, 'v')
    _need(score, 'score', torch.float32)
    _need(stats, 'stats', torch.float32)
    _need(D, 'D', torch.float32)
    grad_v = torch.empty_like(v)
    grad_score = torch.empty_like(score)
    if n_src == 0:
        return grad_v, grad_score
    n_long = 0 if long_ids is None else long_ids.numel()
    with torch.cuda.device(v.device):
        _check(lib().allset_pma_bwd(_ptr(grad_out), _ptr(v), _ptr(score), _ptr(stats), _ptr(D), _dtype_code(v), H, C,
                                    float(slope), _ptr(rowptrT), _ptr(colT), n_src, _ptr(long_ids), n_long,
                                    long_threshold, _ptr(grad_v), _ptr(grad_score), _stream()), 'allset_pma_bwd')
    return grad_v, grad_score
