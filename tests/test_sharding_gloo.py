"""Host-side multi-GPU logic on CPU: partition planner + the in-place ragged all-gather, world_size 2 over gloo.

The compute inside each rank is stood in for by the ORACLE (tests may use it; the product never does): what is
under test is that target-range sharding + the two exchanges reproduce the single-process result exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import allset_oracle as O
from allset_b200 import sharding, synthetic


def _csr(tgt, src, n_tgt):
    order = torch.argsort(tgt, stable=True)
    counts = torch.bincount(tgt, minlength=n_tgt)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)]).int()
    return rowptr, src[order].int()


def test_balanced_ranges_cover_and_balance():
    ei = synthetic.powerlaw_hypergraph(3000, 500, 2, 256, 2.0, seed=3)
    rowptr, col = _csr(ei[1] - 3000, ei[0], 500)
    for parts in (1, 2, 3, 8):
        r = sharding.balanced_ranges(rowptr, parts)
        assert r[0][0] == 0 and r[-1][1] == 500 and all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
        nnz = [int(rowptr[hi]) - int(rowptr[lo]) for lo, hi in r]
        assert sum(nnz) == col.numel()
        assert max(nnz) <= col.numel() / parts + 256 + 1          # off by at most one (max-size) segment
    # one giant segment cannot be split: the other ranks take what is left
    rp = torch.tensor([0, 1000, 1001, 1002, 1003], dtype=torch.int32)
    r = sharding.balanced_ranges(rp, 2)
    assert r == [(0, 1), (1, 4)] or r == [(0, 0), (0, 4)] or r[0][1] <= 1
    assert sharding.balanced_ranges(torch.zeros(1, dtype=torch.int32), 4) == [(0, 0)] * 4


def test_choose_ranges_prefers_equal_rows_when_balanced():
    ei = synthetic.poisson_hypergraph(20000, 4000, 20, seed=5)
    rowptr, _ = _csr(ei[1] - 20000, ei[0], 4000)
    assert sharding.choose_ranges(rowptr, 4, tolerance=0.05) == sharding.equal_ranges(4000, 4)
    skew = torch.cat([torch.zeros(1, dtype=torch.int64), torch.arange(1, 4001).cumsum(0)]).int()
    r = sharding.choose_ranges(skew, 4)
    assert r != sharding.equal_ranges(4000, 4) and r[0][1] > 1000


def test_slice_csr():
    ei = synthetic.poisson_hypergraph(300, 60, 6, seed=7)
    rowptr, col = _csr(ei[1] - 300, ei[0], 60)
    rp, c, p0 = sharding.slice_csr(rowptr, col, 10, 25)
    assert int(rp[0]) == 0 and rp.numel() == 16 and c.numel() == int(rp[-1]) and p0 == int(rowptr[10])
    assert torch.equal(c, col[int(rowptr[10]):int(rowptr[25])])
    with pytest.raises(ValueError):
        sharding.slice_csr(rowptr, col, 5, 100)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ragged, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n, m, d = 400, 90, 8
        ei = (synthetic.powerlaw_hypergraph(n, m, 2, 64, 2.0, seed=11) if ragged
              else synthetic.poisson_hypergraph(n, m, 5, seed=11))
        node, he = ei[0], ei[1] - n
        x_v = torch.randn(n, d)
        rp_e, col_e = _csr(he, node, m)
        rp_v, col_v = _csr(node, he, n)
        pick = sharding.balanced_ranges if ragged else sharding.choose_ranges
        e_ranges, v_ranges = pick(rp_e, world), pick(rp_v, world)

        def local_reduce(src_rows, rowptr, col, lo, hi, out_full):
            rp, c, _ = sharding.slice_csr(rowptr, col, lo, hi)
            tgt = torch.repeat_interleave(torch.arange(hi - lo), (rp[1:] - rp[:-1]).long())
            if hi > lo:
                out_full[lo:hi] = O.scatter_rows(src_rows[c.long()], tgt, 'sum', rows=hi - lo)

        x_e = torch.full((m, d), float('nan'))
        local_reduce(x_v, rp_e, col_e, *e_ranges[rank], x_e)                       # V->E on my hyperedges
        sharding.allgather_rows(x_e, e_ranges, rank)
        x_v2 = torch.full((n, d), float('nan'))
        local_reduce(x_e, rp_v, col_v, *v_ranges[rank], x_v2)                      # E->V on my vertices
        sharding.allgather_rows(x_v2, v_ranges, rank)
        ref_e, ref_v = O.layer_pair_sum(x_v, node, he, None, 'sum')
        ref_v_full = torch.zeros(n, d)
        ref_v_full[:ref_v.shape[0]] = ref_v
        ok = torch.allclose(x_e, ref_e, atol=1e-5) and torch.allclose(x_v2, ref_v_full, atol=1e-5)
        q.put((rank, bool(ok), e_ranges, v_ranges))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('ragged', [False, True])
def test_sharded_layer_pair_world2(ragged):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), res
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]          # every rank plans the same partition


def test_peer_need_mask_matches_brute_force():
    """bit j of row r = the j-th OTHER rank's range contains a neighbour of r (pure index arithmetic, CPU)."""
    g = torch.Generator().manual_seed(4)
    rows, other = 500, 90
    nnz = 4000
    src = torch.sort(torch.randint(0, rows, (nnz,), generator=g))[0]
    col = torch.randint(0, other, (nnz,), generator=g).int()
    rowptr = torch.zeros(rows + 1, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(torch.bincount(src, minlength=rows), 0).int()
    ranges = [(0, 20), (20, 20), (20, 55), (55, 90)]                   # rank 1 owns nothing
    for rank in range(4):
        mask = sharding.peer_need_mask(rowptr, col, rows, ranges, rank)
        others = [q for q in range(4) if q != rank]
        for r in range(0, rows, 7):
            nb = col[int(rowptr[r]):int(rowptr[r + 1])].tolist()
            want = 0
            for j, q in enumerate(others):
                lo, hi = ranges[q]
                if any(lo <= c < hi for c in nb):
                    want |= 1 << j
            assert int(mask[r]) == want, (rank, r)
