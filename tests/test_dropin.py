"""The drop-in `layers` / `models` modules resolve what reference train.py star-imports (dev container only: needs the
reference tree and the third-party shims; nothing here runs on the GPU box)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get('ALLSET_REFERENCE_ROOT', '/root/reference')

CODE = r'''
import sys, warnings
warnings.simplefilter('ignore')
import numpy as np
if not hasattr(np, 'int'): np.int = int
from layers import *
from models import *
import allset_b200
assert SetGNN is allset_b200.SetGNN and HalfNLHconv is allset_b200.HalfNLHconv and PMA is allset_b200.PMA
assert UniGCNII is allset_b200.UniGCNII and UniGCNIIConv is allset_b200.UniGCNIIConv
for name in ('HyperGCN', 'CEGCN', 'CEGAT', 'HCHA', 'HNHN', 'HGNN', 'MLP_model', 'UniGCNII', 'HypergraphConv', 'HNHNConv'):
    assert name in globals(), name
assert HCHA.__module__ == 'allset_b200.baselines' and HNHN.__module__ == 'allset_b200.baselines'
assert HypergraphConv.__module__ == 'allset_b200.baselines' and UniGNN.__module__ == 'allset_b200.baselines'
assert HyperGCN.__module__.startswith('_allset_reference_') and CEGCN.__module__.startswith('_allset_reference_')
from types import SimpleNamespace
args = SimpleNamespace(All_num_layers=1, dropout=0.5, aggregate='add', normalization='ln', deepset_input_norm=True,
                       GPR=False, LearnMask=False, num_features=10, MLP_hidden=8, MLP_num_layers=2, heads=1, PMA=False,
                       Classifier_hidden=8, Classifier_num_layers=1, num_classes=3)
m = SetGNN(args); m.reset_parameters()
print('ok', sum(p.numel() for p in m.parameters()))
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'src', 'train.py')), reason='reference tree not present')
def test_dropin_modules_shadow_the_reference():
    env = dict(os.environ)
    env['ALLSET_REFERENCE_SRC'] = os.path.join(REF, 'src')
    env['PYTHONPATH'] = os.pathsep.join([os.path.join(ROOT, 'allset_b200', 'dropin'), os.path.join(ROOT, 'oracle', 'shims'),
                                         os.path.join(REF, 'src')])
    res = subprocess.run([sys.executable, '-c', CODE], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.strip().startswith('ok')
