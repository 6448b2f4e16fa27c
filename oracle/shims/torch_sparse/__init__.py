"""torch-sparse 0.6.0 `coalesce` restated: sort COO by (row, col), drop duplicates (value=None path only).

Call site: reference load_other_datasets.py:177-181. Test infrastructure only."""
import torch


def coalesce(index, value, m, n, op='add'):
    assert value is None
    key = index[0] * n + index[1]
    key, perm = torch.sort(key, stable=True)
    keep = torch.ones_like(key, dtype=torch.bool)
    keep[1:] = key[1:] != key[:-1]
    return index[:, perm][:, keep], None
