"""Drop-in `layers` module: put this directory BEFORE the reference's `src/` on sys.path (PYTHONPATH=.../allset_b200/dropin)
and `train.py` picks up the B200-native MLP / PMA / HalfNLHconv / HypergraphConv / HNHNConv; the remaining layers (HGNN_conv,
HyperGraphConvolution: dense or Laplacian-matrix products, not incidence-list reduces) come from the reference."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from allset_b200.dropin._forward import load_reference as _load, public_names as _names  # noqa: E402

globals().update(_names(_load('layers')))
from allset_b200.layers import MLP, PMA, HalfNLHconv  # noqa: E402,F401  (reference src/layers.py:42-199,496-656)
from allset_b200.baselines import HypergraphConv, HNHNConv  # noqa: E402,F401  (src/layers.py:233-494, same kernels)
