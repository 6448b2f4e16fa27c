"""Build liballset_b200.so in-tree with nvcc for sm_100a:  python -m allset_b200.build [--force]

The library links only against the CUDA runtime (no torch, no pybind): the Python host reaches it through
ctypes (allset_b200/_lib.py), any other host through include/allset_b200.h.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SOURCES = [os.path.join(_HERE, 'csrc', 'allset_kernels.cu')]
HEADERS = [os.path.join(ROOT, 'include', 'allset_b200.h')] + \
    [os.path.join(_HERE, 'csrc', f) for f in ('mlp_tcgen05.cuh', 'rowop.cuh', 'linear_wgrad.cuh')]
OUTPUT = os.path.join(_HERE, 'liballset_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-diag-suppress', '1444', '-Wno-deprecated-declarations']


def nvcc_path() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def stale() -> bool:
    if not os.path.isfile(OUTPUT):
        return True
    t = os.path.getmtime(OUTPUT)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return OUTPUT
    cmd = [nvcc_path()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'), '-o', OUTPUT] + SOURCES
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed with exit code %d' % res.returncode)
    if verbose:
        sys.stderr.write(res.stderr)
    return OUTPUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
