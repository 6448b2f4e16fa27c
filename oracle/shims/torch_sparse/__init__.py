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


def from_scipy(A):
    """torch-sparse 0.6.0 `from_scipy`: COO (row, col) int64 index + values of a scipy sparse matrix (train.py:401)."""
    A = A.tocoo()
    row, col = torch.from_numpy(A.row).to(torch.long), torch.from_numpy(A.col).to(torch.long)
    return torch.stack([row, col], dim=0), torch.from_numpy(A.data)
