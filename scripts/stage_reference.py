#!/usr/bin/env python
"""Stage the UNMODIFIED reference (its `src/*.py` and the raw cocitation datasets) under `baseline/_ref/AllSet/`.

`baseline/_ref/` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the staged tree
travels to the GPU box -- the same mechanism the bench contract uses for a pip-installed reference.  It exists for ONE
purpose: `tests/test_train_dropin.py` / `scripts/run_train.py` execute the reference's own `train.py`
(reference src/train.py:220-528) end to end with `allset_b200/dropin` in front of it on `sys.path`.  Nothing in
`allset_b200/` reads it.  Run in the dev container (where /root/reference exists):

    python scripts/stage_reference.py [--datasets cora citeseer ...]
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys
import zipfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('ALLSET_REFERENCE_ROOT', '/root/reference')
DEST = os.path.join(ROOT, 'baseline', '_ref', 'AllSet')

# dataset name -> directory inside AllSet_all_raw_data/ (reference src/train.py:312-322)
RAW_DIRS = {'cora': 'cocitation/cora', 'citeseer': 'cocitation/citeseer', 'pubmed': 'cocitation/pubmed',
            'coauthor_cora': 'coauthorship/cora', 'coauthor_dblp': 'coauthorship/dblp', 'zoo': 'zoo',
            'house-committees': 'house-committees', 'NTU2012': 'NTU2012', 'Mushroom': 'Mushroom',
            'walmart-trips': 'walmart-trips', '20newsW100': '20newsW100', 'ModelNet40': 'ModelNet40', 'yelp': 'yelp'}


def stage(datasets) -> str:
    src = os.path.join(REF, 'src')
    if not os.path.isfile(os.path.join(src, 'train.py')):
        raise SystemExit('reference tree not found at %s' % REF)
    os.makedirs(os.path.join(DEST, 'src'), exist_ok=True)
    for f in sorted(os.listdir(src)):
        if f.endswith('.py'):
            shutil.copyfile(os.path.join(src, f), os.path.join(DEST, 'src', f))
    raw_zip = os.path.join(REF, 'data', 'raw_data', 'AllSet_all_raw_data.zip')
    with zipfile.ZipFile(raw_zip) as z:
        for name in datasets:
            prefix = 'AllSet_all_raw_data/%s/' % RAW_DIRS[name]
            for info in z.infolist():
                if info.filename.startswith(prefix) and not info.is_dir() and '/splits/' not in info.filename \
                        and '.ipynb' not in info.filename:
                    out = os.path.join(DEST, 'data', info.filename)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    with open(out, 'wb') as fh:
                        fh.write(z.read(info))
    return DEST


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--datasets', nargs='*', default=['cora', 'citeseer'])
    a = ap.parse_args()
    print(stage(a.datasets))
    sys.exit(0)
