"""PyG 1.6.3 `MessagePassing.propagate` restated for Tensor `edge_index`.

flow='source_to_target': j = edge_index[0] (source), i = edge_index[1] (target); 'target_to_source' swaps the rows.
`size` = (N_j, N_i) for either flow, dim_size = N_i (1.6.3 `__collect__`).  Arguments of message()/
aggregate()/update() are resolved by NAME: `<k>_j` / `<k>_i` -> index_select of kwargs[k] along node_dim,
specials `index, ptr, size_i, size_j, dim_size, edge_index*`, everything else passed through from kwargs.
Test infrastructure only.
"""
import inspect
import torch
from torch_scatter import scatter


class MessagePassing(torch.nn.Module):
    special_args = {'edge_index', 'edge_index_i', 'edge_index_j', 'size', 'size_i', 'size_j',
                    'index', 'ptr', 'dim_size', 'adj_t'}

    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert flow in ('source_to_target', 'target_to_source')

    @staticmethod
    def _params(fn, skip_first):
        ps = list(inspect.signature(fn).parameters.values())
        if skip_first:
            ps = ps[1:]
        return ps

    def propagate(self, edge_index, size=None, **kwargs):
        i, j = (1, 0) if self.flow == 'source_to_target' else (0, 1)
        size = [None, None] if size is None else list(size)
        idx = {'_i': i, '_j': j}

        def lift(name):
            # PyG 1.6.3 __collect__: `size` is (N_j, N_i) -- slot 0 belongs to the `_j` (source) set and slot 1 to the
            # `_i` (target) set WHATEVER the flow; only the row of edge_index that is lifted depends on the flow.
            for suf in ('_i', '_j'):
                if name.endswith(suf):
                    slot = 0 if suf == '_j' else 1
                    data = kwargs.get(name[:-2], None)
                    if isinstance(data, (tuple, list)):
                        data = data[slot]
                    if isinstance(data, torch.Tensor):
                        if size[slot] is None:
                            size[slot] = data.size(self.node_dim)
                        return True, data.index_select(self.node_dim, edge_index[idx[suf]])
                    return True, data
            return False, None

        wanted = {}
        for fn, skip in ((self.message, False), (self.aggregate, True), (self.update, True)):
            for p in self._params(fn, skip):
                if p.name in wanted or p.name in self.special_args:
                    continue
                hit, val = lift(p.name)
                if hit:
                    wanted[p.name] = val
                elif p.name in kwargs:
                    wanted[p.name] = kwargs[p.name]
                elif p.default is not inspect.Parameter.empty:
                    wanted[p.name] = p.default
                else:
                    raise TypeError('missing argument %r for propagate' % p.name)
        size_i = size[1] if size[1] is not None else size[0]
        size_j = size[0] if size[0] is not None else size[1]
        wanted.update(edge_index=edge_index, edge_index_i=edge_index[i], edge_index_j=edge_index[j],
                      index=edge_index[i], ptr=None, size=size, size_i=size_i, size_j=size_j,
                      dim_size=size_i, adj_t=None)

        def call(fn, first, skip):
            kw = {p.name: wanted[p.name] for p in self._params(fn, skip)}
            return fn(*first, **kw)

        out = call(self.message, (), False)
        out = call(self.aggregate, (out,), True)
        return call(self.update, (out,), True)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def update(self, inputs):
        return inputs
