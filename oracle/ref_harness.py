"""Import the reference's OWN modules, unmodified, in the dev container.  TEST INFRASTRUCTURE ONLY.

`/root/reference` exists only in the dev container (never on the GPU box), and its third-party dependencies
(torch_scatter, torch_geometric, torch_sparse, ipdb, matplotlib) are not installed, so this harness
  1. puts `oracle/shims/` (published-semantics stand-ins, see its README) and `/root/reference/src` on sys.path,
  2. restores `np.int` (removed in numpy >= 1.24; used at load_other_datasets.py:166),
  3. imports the reference's `layers`, `models`, `preprocessing`, `load_other_datasets` modules as they are.
Used by the `oracle/make_golden*.py` fixture generators and by `tests/test_ingest.py` (skipped when the reference tree
is absent).  Nothing here is reachable from `allset_b200/`.
"""
from __future__ import annotations

import importlib
import io
import os
import sys
import tempfile
import warnings
import zipfile

REFERENCE_ROOT = os.environ.get('ALLSET_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')
_RAW_ZIP = os.path.join(REFERENCE_ROOT, 'data', 'raw_data', 'AllSet_all_raw_data.zip')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'layers.py'))


def load():
    """Returns a namespace with the reference modules: .layers .models .preprocessing .loaders"""
    if not available():
        raise RuntimeError('reference tree not found at %s' % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int                      # noqa: the reference predates numpy 1.24
    for p in (os.path.join(REFERENCE_ROOT, 'src'), _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from types import SimpleNamespace
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')   # `is 'Identity'` SyntaxWarning at layers.py:568,612,614
        mods = SimpleNamespace(
            layers=importlib.import_module('layers'),
            models=importlib.import_module('models'),
            preprocessing=importlib.import_module('preprocessing'),
            loaders=importlib.import_module('load_other_datasets'),
        )
    assert mods.layers.__file__.startswith(REFERENCE_ROOT), mods.layers.__file__
    assert mods.models.__file__.startswith(REFERENCE_ROOT), mods.models.__file__
    return mods


def load_cocitation(name: str):
    """Run the reference's own pipeline for a cocitation dataset (train.py:308-353 with default flags):
    load_citation_dataset -> ExtractV2E -> Add_Self_Loops -> norm_contruction('all_one').

    Returns the reference `Data` object: x [N,F] f32, edge_index [2,nnz] i64 (row0 node ascending, row1 hyperedge
    id in [N, N+M)), norm [nnz] i64 ones, y [N]."""
    import torch
    mods = load()
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(_RAW_ZIP) as z:
        base = 'AllSet_all_raw_data/cocitation/%s/' % name
        os.makedirs(os.path.join(tmp, name))
        for f in ('features.pickle', 'labels.pickle', 'hypergraph.pickle'):
            with open(os.path.join(tmp, name, f), 'wb') as out:
                out.write(z.read(base + f))
        with warnings.catch_warnings(), _quiet():
            warnings.simplefilter('ignore')
            data = mods.loaders.load_citation_dataset(path=tmp, dataset=name)
    # train.py:334-339 wraps these in tensors when they are missing; the loaders return python ints, and
    # ExtractV2E / Add_Self_Loops index them with [0] (preprocessing.py:400-401,416-417).
    data.n_x = torch.tensor([data.n_x])
    data.num_hyperedges = torch.tensor([data.num_hyperedges])
    data = mods.preprocessing.ExtractV2E(data)
    data = mods.preprocessing.Add_Self_Loops(data)
    data = mods.preprocessing.norm_contruction(data, option='all_one')
    return data


class _quiet(object):
    def __enter__(self):
        self._o = sys.stdout
        sys.stdout = io.StringIO()

    def __exit__(self, *a):
        sys.stdout = self._o
