"""The reference's UNMODIFIED `train.py` runs end to end on a B200 with `allset_b200/dropin` in front of it
(reference src/train.py:28-42,437,462,478 -- north_star: "drops into src/train.py unchanged").

The reference tree comes from `baseline/_ref/AllSet` (staged by scripts/stage_reference.py in the dev container;
git-ignored, travels to the GPU box with the snapshot).  Skipped when it is absent.  `scripts/run_train.py` does the
sys.path / shim plumbing and reports what train.py appended to its own results CSV.
"""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

STAGED = os.path.join(ROOT, 'baseline', '_ref', 'AllSet', 'src', 'train.py')
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isfile(STAGED), reason='reference not staged')]

CORA = ['--method', 'AllDeepSets', '--dname', 'cora', '--All_num_layers', '1', '--MLP_num_layers', '2',
        '--Classifier_num_layers', '1', '--MLP_hidden', '64', '--Classifier_hidden', '64', '--wd', '0',
        '--feature_noise', '0.0', '--cuda', '0', '--lr', '0.01']     # run_one_model.sh:39-55 (lr raised: 80 epochs here)
CITESEER = ['--method', 'AllSetTransformer', '--dname', 'citeseer', '--All_num_layers', '2', '--MLP_num_layers', '2',
            '--Classifier_num_layers', '1', '--MLP_hidden', '128', '--Classifier_hidden', '128', '--heads', '4',
            '--wd', '0', '--feature_noise', '0.0', '--cuda', '0', '--lr', '0.01']    # BASELINE.json configs[1]


def _run(train_args, agg_dtype=None, epochs=80, runs=3):
    cmd = [sys.executable, os.path.join(ROOT, 'scripts', 'run_train.py'), '--impl', 'dropin']
    if agg_dtype:
        cmd += ['--agg-dtype', agg_dtype]
    cmd += ['--'] + train_args + ['--epochs', str(epochs), '--runs', str(runs)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith('{')][-1]
    return json.loads(line)


@pytest.mark.parametrize('agg_dtype', [None, 'bf16'])
def test_train_py_cora_alldeepsets(agg_dtype):
    r = _run(CORA, agg_dtype)
    assert r['setgnn_class'] == 'allset_b200.models' and r['models_module'].endswith('dropin/models.py')
    assert any(p.endswith('liballset_b200.so') for p in r['native_so_loaded'])
    assert r['device'] != 'cpu' and r['params'] == 125369            # parameter count of the reference model
    # the reference's own CPU run of these flags ends at 52 % (lr 1e-3, 500 epochs; profiles/r02_train_py.md) with a 5-10
    # point spread over splits on every implementation: 80 epochs at lr 1e-2 must be clearly above chance (1/7 = 14 %)
    assert r['best_val_acc_mean'] > 25.0 and r['test_acc_mean'] > 22.0, r


def test_train_py_citeseer_allsettransformer():
    r = _run(CITESEER, None, epochs=60)
    assert r['setgnn_class'] == 'allset_b200.models'
    assert any(p.endswith('liballset_b200.so') for p in r['native_so_loaded'])
    assert r['test_acc_mean'] > 40.0, r                              # chance = 1/6; the reference ends at 72.6 %
