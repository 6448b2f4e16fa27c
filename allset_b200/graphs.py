"""CUDA-graph replay of a SetGNN forward for launch-bound graphs.

On the real datasets (cora: 2708 nodes, 7494 incidences) one forward is ~40 kernel launches of a few microseconds
each, so the step is bound by launch latency, not by the device: capture the whole forward once and replay it.
Everything on the path is capturable: the aggregation kernels enqueue on the current stream through the C ABI, the
incidence CSR is cached on `data.edge_index` after the first call, and nothing synchronises with the host.
"""
from __future__ import annotations

import torch


class GraphedForward(object):
    """`g = GraphedForward(model, data); out = g()` replays `model(data)` (inference: eval mode, no autograd).
    `data.x` may be refreshed in place (`g.x.copy_(new_x)`) between replays; the graph structure is fixed."""

    def __init__(self, model, data, warmup: int = 3):
        if not data.x.is_cuda:
            raise RuntimeError('allset_b200.GraphedForward needs CUDA tensors (no CPU path)')
        self.model, self.data = model, data
        self.x = data.x
        model.eval()
        side = torch.cuda.Stream(device=data.x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                 # builds + caches the incidence, opts kernels into their smem
                model(data)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = model(data)

    def __call__(self):
        self.graph.replay()
        return self.out
