"""pytest configuration: the `gpu` marker, import paths, golden-fixture loaders.

`-m "not gpu"` runs on the CPU-only dev container (oracle vs golden vectors, host logic, C-ABI symbol check,
gloo sharding); `-m gpu` runs the parity tests proper on a B200 through the C ABI.  Only tests/ (and bench.py's
cpu_baseline leg, __graft_entry__.smoke) may import oracle/.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


_cache = {}


def load_golden(name):
    if name not in _cache:
        _cache[name] = torch.load(os.path.join(GOLDEN, name), map_location='cpu', weights_only=True)
    return _cache[name]


def densify(sp):
    x = torch.zeros(sp['shape'], dtype=sp['vals'].dtype)
    x[sp['rows'].long(), sp['cols'].long()] = sp['vals']
    return x


def golden_x(rec):
    return densify(rec['x_sparse']) if 'x_sparse' in rec else rec['x']


@pytest.fixture(scope='session')
def golden():
    return load_golden
