"""Minimal `Data` / `InMemoryDataset` stand-ins so the reference's loaders import. Test infrastructure only."""


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.size(0)


class InMemoryDataset(object):
    def __init__(self, *a, **k):
        raise NotImplementedError('dataset caching is out of scope for the oracle')
