"""allset_b200 -- B200-native implementation of AllSet's V->E / E->V multiset-aggregation path.

Public surface (mirrors the reference's module API, reference src/layers.py + src/models.py):
    SetGNN, HalfNLHconv, PMA, MLP          drop-in modules (same ctor / forward / state_dict)
    Incidence                               the sorted incidence container (CSR by target + CSR by source)
    segment_reduce, pma_aggregate           the two differentiable aggregation operators
    UniGCNII, UniGCNIIConv                  the reference's UniGCNII baseline on the same kernels (V->E mean, E->V sum)
    HypergraphConv/HCHA, HNHNConv/HNHN, UniGNN (+ UniSAGE/GIN/GCN/GAT convs)   the other incidence-list baselines (baselines.py)
    GraphedForward                          CUDA-graph replay of a SetGNN forward (launch-bound real datasets)
    preprocessing, sharding, synthetic      incidence preprocessing, multi-GPU partition + fused exchange, generators
    ingest                                  star expansion of a hyperedge dictionary + flat memory-mapped dataset cache
The device code lives in liballset_b200.so (C ABI: include/allset_b200.h), built by `python -m allset_b200.build`.
There is no CPU fallback anywhere in this package.
"""
from .graph import Incidence, Csr, incidence_of  # noqa: F401
from .ops import segment_reduce, pma_aggregate  # noqa: F401
from .layers import MLP, PMA, HalfNLHconv  # noqa: F401
from .models import SetGNN  # noqa: F401
from .graphs import GraphedForward  # noqa: F401
from .uni import UniGCNII, UniGCNIIConv  # noqa: F401
from . import baselines  # noqa: F401
from .baselines import HypergraphConv, HCHA, HNHNConv, HNHN, UniGNN  # noqa: F401

__version__ = '0.1.0'
