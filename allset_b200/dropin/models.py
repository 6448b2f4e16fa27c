"""Drop-in `models` module: `SetGNN` is the B200-native one (same ctor / forward / state_dict as reference
src/models.py:295-484); `UniGCNII` / `UniGCNIIConv` (reference src/models.py:909-995) run on the same segmented-reduce kernels; the other
the HCHA / HGNN, HNHN and UniGNN-family baselines (reference src/models.py:207-292,601-907) likewise; the remaining baseline
models (HyperGCN, CEGCN / CEGAT: clique-expansion GCNs, no incidence-list reduce) are forwarded from the reference."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from allset_b200.dropin._forward import load_reference as _load, public_names as _names  # noqa: E402

globals().update(_names(_load('models')))
from allset_b200.models import SetGNN as _SetGNN  # noqa: E402

_AGG = {'bf16': 'bfloat16', 'f32': 'float32'}.get(_os.environ.get('ALLSET_AGG_DTYPE', ''))
if _AGG is None:
    SetGNN = _SetGNN
else:
    # train.py constructs `SetGNN(args)` / `SetGNN(args, data.norm)` (reference src/train.py:32-42) and has no flag for
    # the storage dtype of the gathered rows: ALLSET_AGG_DTYPE=bf16 selects the bf16 mode without touching train.py
    import torch as _torch

    class SetGNN(_SetGNN):
        def __init__(self, args, norm=None, agg_dtype=getattr(_torch, _AGG)):
            super().__init__(args, norm, agg_dtype=agg_dtype)
from allset_b200.uni import UniGCNII, UniGCNIIConv  # noqa: E402,F401  (same kernels)
from allset_b200.baselines import (HCHA, HNHN, UniGNN, UniSAGEConv, UniGINConv, UniGCNConv, UniGCNConv2,  # noqa: E402,F401
                                   UniGATConv)
