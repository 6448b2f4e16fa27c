"""Fixtures for allset_b200.preprocessing.expand_edge_index from the reference's OWN expand_edge_index
(reference src/preprocessing.py:22-144), run unmodified under oracle/ref_harness.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden_expand.py    ->  tests/golden/expand_edge_index.pt

Cases: real cora after ExtractV2E + Add_Self_Loops (the train.py:344-349 sequence with --exclude_self), random
hypergraphs with singleton hyperedges and nodes of degree 1, the edge_th trimming option, a list without self loops.
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'expand_edge_index.pt')


def run(mods, name, ei, n_x, n_he, totedges=None, edge_th=0):
    d = SimpleNamespace(edge_index=ei.clone(), n_x=torch.tensor([n_x]), num_hyperedges=torch.tensor([n_he]))
    if totedges is not None:
        d.totedges = totedges
    d = mods.preprocessing.expand_edge_index(d, edge_th=edge_th)
    return {'name': name, 'edge_index': ei.clone(), 'n_x': n_x, 'n_he': totedges if totedges is not None else n_he,
            'edge_th': edge_th, 'expanded': d.edge_index.clone()}


def random_v2e(n, m, max_size, seed):
    g = torch.Generator().manual_seed(seed)
    nodes, hes = [], []
    for e in range(m):
        s = int(torch.randint(1, max_size + 1, (1,), generator=g))
        members = torch.randperm(n, generator=g)[:s]
        nodes.append(members)
        hes.append(torch.full((s,), n + e, dtype=torch.long))
    ei = torch.stack([torch.cat(nodes), torch.cat(hes)])
    return ei[:, torch.sort(ei[0], stable=True)[1]]


def main():
    mods = ref_harness.load()
    cases = []
    with ref_harness._quiet():
        data = ref_harness.load_cocitation('cora')
    cases.append(run(mods, 'cora + self loops', data.edge_index, 2708, 1579, totedges=int(data.totedges)))
    cases.append(run(mods, 'random 60x40', random_v2e(60, 40, 6, 1), 60, 40))
    cases.append(run(mods, 'random 30x25 edge_th=3', random_v2e(30, 25, 7, 2), 30, 25, edge_th=3))
    cases.append(run(mods, 'random 200x50 big', random_v2e(200, 50, 25, 3), 200, 50))
    cases.append(run(mods, 'all singletons', torch.stack([torch.arange(12), 12 + torch.arange(12)]), 12, 12))
    torch.save(cases, OUT)
    for c in cases:
        print(c['name'], tuple(c['edge_index'].shape), '->', tuple(c['expanded'].shape))
    print('%s %.2f MB' % (OUT, os.path.getsize(OUT) / 1e6))


if __name__ == '__main__':
    main()
