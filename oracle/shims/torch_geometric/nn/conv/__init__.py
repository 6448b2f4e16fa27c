from .message_passing import MessagePassing


class _Unavailable(MessagePassing):
    """GCNConv/GATConv are only used by out-of-scope baselines (reference models.py:80-183)."""

    def __init__(self, *a, **k):
        raise NotImplementedError('baseline conv not provided by the oracle stand-in')


GCNConv = _Unavailable
GATConv = _Unavailable
