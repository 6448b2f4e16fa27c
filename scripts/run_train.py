#!/usr/bin/env python
"""Execute the reference's UNMODIFIED `train.py` (reference src/train.py:220-528) end to end. TEST TOOLING.

    python scripts/run_train.py --impl dropin    -- --method AllDeepSets --dname cora --cuda 0 ...   # B200-native SetGNN
    python scripts/run_train.py --impl reference -- --method AllDeepSets --dname cora --cuda -1 ...  # reference modules

`--impl dropin` puts `allset_b200/dropin` in front of the reference's `src/` on sys.path, so `from layers import *` /
`from models import *` (train.py:21-22) pick up the B200-native SetGNN / HalfNLHconv / PMA / MLP and everything else
(argument parsing, dataset cache, preprocessing, split, Adam loop, Logger, CSV) is the reference's own code.
`--impl reference` runs the same file against the reference's own modules (CPU with `--cuda -1`, or the GPU through
ATen `scatter_add_` with `--cuda 0`): the accuracy / seconds-per-run baseline beside it.

The reference tree comes from `baseline/_ref/AllSet` (scripts/stage_reference.py; git-ignored) or, in the dev
container, /root/reference.  Third-party packages the reference pins and this image lacks (torch_scatter, torch_sparse,
torch_geometric, ipdb, matplotlib) are satisfied by `oracle/shims` -- which is why this lives with the test tooling and
not in the package.  Prints one JSON line with the numbers train.py itself appended to its results CSV.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import shutil
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'baseline', '_ref', 'AllSet')


def reference_root(workdir: str) -> str:
    """A WRITABLE copy of the reference layout (train.py writes ../data/pyg_data and ./hyperparameter_tunning)."""
    if os.path.isfile(os.path.join(STAGED, 'src', 'train.py')):
        return STAGED
    ref = os.environ.get('ALLSET_REFERENCE_ROOT', '/root/reference')
    if not os.path.isfile(os.path.join(ref, 'src', 'train.py')):
        raise SystemExit('no reference tree: run scripts/stage_reference.py in the dev container first')
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    import stage_reference
    return stage_reference.stage(['cora', 'citeseer'])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--impl', default='dropin', choices=['dropin', 'reference'])
    ap.add_argument('--agg-dtype', default=None, choices=[None, 'bf16', 'f32'],
                    help='dropin only: storage dtype of the gathered rows (ALLSET_AGG_DTYPE for dropin/models.py)')
    ap.add_argument('--fresh-cache', action='store_true', help='delete the processed dataset cache first')
    ap.add_argument('train_args', nargs=argparse.REMAINDER)
    a = ap.parse_args()
    targs = [t for t in a.train_args if t != '--']
    ref = reference_root(ROOT)
    src = os.path.join(ref, 'src')
    os.environ['ALLSET_REFERENCE_SRC'] = src
    if a.agg_dtype:
        os.environ['ALLSET_AGG_DTYPE'] = a.agg_dtype
    paths = [os.path.join(ROOT, 'oracle', 'shims'), src]
    if a.impl == 'dropin':
        paths.insert(0, os.path.join(ROOT, 'allset_b200', 'dropin'))
    sys.path[:0] = paths
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int                       # noqa: the reference predates numpy 1.24 (load_other_datasets.py:166)
    if a.fresh_cache:
        shutil.rmtree(os.path.join(ref, 'data', 'pyg_data'), ignore_errors=True)
    os.chdir(src)

    def flag(name, default=None):
        return targs[targs.index(name) + 1] if name in targs else default

    dname, noise = flag('--dname', 'walmart-trips-100'), flag('--feature_noise', '1')
    csv = os.path.join(src, 'hyperparameter_tunning', '%s_noise_%s.csv' % (dname, noise))
    before = sum(1 for _ in open(csv)) if os.path.isfile(csv) else 0
    sys.argv = ['train.py'] + targs
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            runpy.run_path(os.path.join(src, 'train.py'), run_name='__main__')
        except SystemExit as e:            # train.py ends with quit()
            if e.code not in (None, 0):
                raise
    wall = time.time() - t0
    lines = open(csv).read().splitlines()
    assert len(lines) == before + 1, 'train.py did not append its result line to %s' % csv
    f = [t.strip() for t in lines[-1].split(',')]
    # '<method>_<lr>_<wd>_<heads>, val mean ± std, test mean ± std, params, avg s, std s, XminYs'  (train.py:511-517)
    val, test = [tuple(float(v) for v in t.split('±')) for t in f[1:3]]
    import torch
    loaded = sorted({ln.split()[-1] for ln in open('/proc/self/maps') if 'allset_b200' in ln and ln.rstrip().endswith('.so')})
    dev = 'cpu' if flag('--cuda', '0') == '-1' or not torch.cuda.is_available() else torch.cuda.get_device_name(0)
    print(json.dumps({
        'impl': a.impl, 'agg_dtype': a.agg_dtype, 'method': flag('--method', 'AllSetTransformer'), 'dname': dname,
        'device': dev, 'runs': int(flag('--runs', 20)), 'epochs': int(flag('--epochs', 500)),
        'best_val_acc_mean': val[0], 'best_val_acc_std': val[1], 'test_acc_mean': test[0], 'test_acc_std': test[1],
        'params': int(f[3]), 'seconds_per_run_mean': float(f[4].rstrip('s')), 'seconds_per_run_std': float(f[5].rstrip('s')),
        'wall_s': wall, 'train_args': targs, 'native_so_loaded': loaded,
        'models_module': sys.modules['models'].__file__,
        'setgnn_class': next((c.__module__ for c in sys.modules['models'].SetGNN.__mro__ if c.__module__.startswith('allset_b200')),
                             sys.modules['models'].SetGNN.__module__),
    }), flush=True)


if __name__ == '__main__':
    main()
