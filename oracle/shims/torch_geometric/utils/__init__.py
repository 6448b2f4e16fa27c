"""PyG 1.6.3 `utils.softmax` restated; call site: reference layers.py:174."""
from torch_scatter import scatter


def softmax(src, index, ptr=None, num_nodes=None):
    assert ptr is None, 'this stand-in only covers the index path the reference uses'
    N = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = src - scatter(src, index, dim=0, dim_size=N, reduce='max')[index]
    out = out.exp()
    out_sum = scatter(out, index, dim=0, dim_size=N, reduce='sum')[index]
    return out / (out_sum + 1e-16)


def degree(index, num_nodes=None, dtype=None):
    import torch
    N = int(index.max()) + 1 if num_nodes is None else int(num_nodes)
    out = torch.zeros((N,), dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones((index.size(0),)))
