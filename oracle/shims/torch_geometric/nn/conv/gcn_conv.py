def gcn_norm(*a, **k):
    raise NotImplementedError('gcn_norm is only used by the out-of-scope CEGCN/CEGAT preprocessing')
