"""Stand-in for torch-scatter 2.0.4 (reference README.md:18-22), CPU/ATen only.

Restates the semantics the reference relies on (call sites: layers.py:194,656; preprocessing.py:459-460):
  * `scatter_sum` = zeros(size).scatter_add_(dim, broadcast(index), src); size[dim] = dim_size or index.max()+1
  * `scatter_mean` = scatter_sum / clamp(count, min=1)
  * `scatter(..., reduce=)` dispatch; 'min'/'max' leave empty segments at 0.
Test infrastructure only (see oracle/shims/README.md).
"""
import torch


def broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand_as(src)


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = dim_size
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count.clamp_(1)
    count = broadcast(count, out, dim)
    if torch.is_floating_point(out):
        out.div_(count)
    else:
        out.floor_divide_(count)
    return out


def _scatter_minmax(src, index, dim, out, dim_size, which):
    index = broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(dim, index, src, reduce=which, include_self=False)


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_minmax(src, index, dim, out, dim_size, 'amax'), None


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_minmax(src, index, dim, out, dim_size, 'amin'), None


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce='sum'):
    if reduce in ('sum', 'add'):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == 'mean':
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == 'min':
        return scatter_min(src, index, dim, out, dim_size)[0]
    if reduce == 'max':
        return scatter_max(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)
