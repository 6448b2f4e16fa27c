"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol the header
declares, the modules mirror the reference's constructor / state_dict surface, and the product path fails loudly
instead of falling back to any CPU implementation."""
import ctypes
import os
import re
from types import SimpleNamespace

import pytest
import torch

import allset_oracle as O
from conftest import ROOT, load_golden


def _header_functions():
    text = open(os.path.join(ROOT, 'include', 'allset_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(allset_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def built_lib():
    from allset_b200 import build
    return build.build()


def test_header_declares_the_documented_entry_points():
    names = _header_functions()
    for must in ('allset_version', 'allset_last_error', 'allset_csr_from_coo', 'allset_segreduce_fwd',
                 'allset_segreduce_bwd_w', 'allset_pma_fwd', 'allset_pma_bwd', 'allset_pma_alpha'):
        assert must in names


def test_library_exports_every_header_symbol(built_lib):
    h = ctypes.CDLL(built_lib)
    for name in _header_functions():
        assert hasattr(h, name), 'liballset_b200.so does not export %s' % name
    h.allset_version.restype = ctypes.c_int
    m = re.search(r'#define ALLSET_ABI_VERSION (\d+)', open(os.path.join(ROOT, 'include', 'allset_b200.h')).read())
    assert h.allset_version() == int(m.group(1))


def test_ctypes_signatures_cover_the_header(built_lib):
    from allset_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_functions()
    assert _lib.lib().allset_version() == _lib.ABI_VERSION


def _header_prototypes():
    """name -> list of parameter declarations, parsed from the header (comments stripped)."""
    text = open(os.path.join(ROOT, 'include', 'allset_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(allset_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', text, flags=re.S):
        params = [q.strip() for q in m.group(2).split(',')]
        protos[m.group(1)] = [] if params in ([''], ['void']) else params
    return protos


def test_ctypes_signatures_match_the_header_prototypes():
    """Arity and the pointer / integer / float kind of every parameter: a ctypes binding that drifts from the header
    (an argument added on one side only) corrupts the call silently."""
    from allset_b200 import _lib
    protos = _header_prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, params in protos.items():
        argtypes = _lib.SIGNATURES[name][1]
        assert len(argtypes) == len(params), '%s: header has %d parameters, ctypes %d' % (name, len(params), len(argtypes))
        for decl, ct in zip(params, argtypes):
            if '*' in decl:
                assert ct in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(ct, 'contents') or ct.__name__.startswith('LP_'), \
                    '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\bfloat\b', decl):
                assert ct is ctypes.c_float, '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\buint64_t\b', decl):
                assert ct is ctypes.c_uint64, '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\bint64_t\b', decl):
                assert ct is ctypes.c_int64, '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\bint32_t\b', decl):
                assert ct is ctypes.c_int32, '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\bsize_t\b', decl):
                assert ct is ctypes.c_size_t, '%s: %r bound as %s' % (name, decl, ct)
            elif re.search(r'\bint\b', decl):
                assert ct is ctypes.c_int, '%s: %r bound as %s' % (name, decl, ct)
            else:
                raise AssertionError('%s: unrecognised parameter %r' % (name, decl))


def test_library_is_sm100a_only(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.isfile(cuobjdump):
        pytest.skip('cuobjdump not available')
    out = subprocess.run([cuobjdump, '--list-elf', built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r'sm_(\d+a?)', out))
    assert archs == {'100a'}, archs


@pytest.mark.parametrize('name', ['cora_alldeepsets.pt', 'citeseer_allsettransformer.pt'])
def test_state_dict_keys_match_reference(name):
    import allset_b200
    rec = load_golden(name)
    model = allset_b200.SetGNN(SimpleNamespace(**rec['args']))
    sd = model.state_dict()
    assert list(sd.keys()) == list(rec['state_dict'].keys())
    for k, v in rec['state_dict'].items():
        assert sd[k].shape == v.shape and sd[k].dtype == v.dtype, k
    model.load_state_dict(rec['state_dict'], strict=True)


def test_state_dict_keys_match_reference_variants():
    import allset_b200
    for rec in load_golden('setgnn_variants.pt'):
        args = SimpleNamespace(**rec['args'])
        model = allset_b200.SetGNN(args, rec['norm']) if args.LearnMask else allset_b200.SetGNN(args)
        assert list(model.state_dict().keys()) == list(rec['state_dict'].keys()), rec['name']
        model.load_state_dict(rec['state_dict'], strict=True)
        model.reset_parameters()
        n_ref = sum(v.numel() for k, v in rec['state_dict'].items() if 'running_' not in k and 'num_batches' not in k)
        assert sum(p.numel() for p in model.parameters()) == n_ref


def test_reset_parameters_distributions():
    """glorot on lin_K / lin_V weights, xavier_uniform on the seed (reference layers.py:96-104)."""
    import math
    import allset_b200
    torch.manual_seed(0)
    p = allset_b200.PMA(64, 128, 128, 2, heads=4)
    bound = math.sqrt(6.0 / (64 + 128))
    assert float(p.lin_K.weight.abs().max()) <= bound and float(p.lin_K.weight.abs().max()) > 0.9 * bound
    assert p.att_r.shape == (1, 4, 32)
    fan_in, fan_out = 4 * 32, 1 * 32           # torch's fan computation for a [1, H, C] tensor
    b2 = math.sqrt(6.0 / (fan_in + fan_out))
    assert float(p.att_r.abs().max()) <= b2
    assert p.rFF.lins[0].in_features == 128 and isinstance(p.rFF.normalizations[0], torch.nn.Identity)


def test_no_cpu_fallback():
    import allset_b200
    rec = load_golden('setgnn_variants.pt')[0]
    model = allset_b200.SetGNN(SimpleNamespace(**rec['args'])).eval()
    data = SimpleNamespace(x=rec['x'], edge_index=rec['edge_index'].clone(), norm=rec['norm'])
    with pytest.raises(RuntimeError, match='CUDA'):
        model(data)
    with pytest.raises(RuntimeError, match='CUDA|CPU'):
        allset_b200.Incidence.from_coo(rec['edge_index'][0], rec['edge_index'][1])
    conv = allset_b200.HalfNLHconv(14, 16, 16, 2, 0.0, 'ln', True, heads=1, attention=False)
    with pytest.raises(RuntimeError, match='CUDA'):
        conv(rec['x'], rec['edge_index'], rec['norm'], 'add')


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'allset_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'allset_oracle' not in text and 'ref_harness' not in text, f
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f


def test_missing_library_fails_loudly(monkeypatch):
    from allset_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/liballset_b200.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.lib()


def test_synthetic_generators_follow_reference_layout():
    from allset_b200 import synthetic
    ei = synthetic.poisson_hypergraph(5000, 1000, 20, seed=1)
    assert ei.dtype == torch.int64 and ei.shape[0] == 2
    assert bool((ei[0, 1:] >= ei[0, :-1]).all())                       # ExtractV2E: sorted by node
    assert int(ei[1].min()) == 5000 and int(ei[1].max()) == 5999       # hyperedge ids start at N
    sizes = torch.bincount(ei[1] - 5000)
    assert int(sizes.min()) >= 1 and abs(float(sizes.float().mean()) - 20) < 0.5
    ei2 = synthetic.poisson_hypergraph(5000, 1000, 20, seed=1)
    assert torch.equal(ei, ei2)
    pl = synthetic.powerlaw_hypergraph(20000, 3000, 2, 4096, 2.0, seed=1)
    s = torch.bincount(pl[1] - 20000)
    assert int(s.max()) == 4096 and int(s.min()) >= 2
    assert synthetic.algorithmic_bytes(4_000_000, 200_000, 128, 2) == 4_000_000 * 260 + 200_001 * 4 + 200_000 * 256
    # the oracle accepts the layout as is
    x = torch.randn(5000, 4)
    out = O.aggregate_sum_mean(x, ei[0], ei[1] - 5000, None, 'sum')
    assert out.shape == (1000, 4)


def test_packed_pma_record_views_layout_cpu():
    """[values | scores] records: both views alias one buffer with the record size as row pitch (no CUDA needed)."""
    from allset_b200 import _lib
    buf, v, s = _lib.packed_pma_records(10, 128, 8, torch.bfloat16, 'cpu')
    assert buf.shape == (10, 128 * 2 + 8 * 4) and buf.dtype == torch.uint8
    assert v.shape == (10, 128) and v.dtype == torch.bfloat16 and v.stride() == (144, 1)
    assert s.shape == (10, 8) and s.dtype == torch.float32 and s.stride() == (72, 1)
    v.fill_(1.0)
    s.fill_(2.0)
    assert v.untyped_storage().data_ptr() == buf.untyped_storage().data_ptr() == s.untyped_storage().data_ptr()
    assert s.data_ptr() - v.data_ptr() == 256 and float(v.float().sum()) == 1280.0 and float(s.sum()) == 160.0
    buf32, v32, s32 = _lib.packed_pma_records(3, 64, 4, torch.float32, 'cpu')
    assert buf32.shape == (3, 64 * 4 + 16) and v32.stride() == (68, 1) and s32.stride() == (68, 1)


def test_tensor_core_paths_are_cuda_eval_bf16_only():
    """The tcgen05 paths need CUDA rows, no_grad, bf16 mode and a square 64/128-wide two-layer MLP; anything else takes
    the fp32 path (or raises for non-CUDA inputs further down) -- never a silent CPU computation of the product path."""
    import allset_b200
    m = allset_b200.MLP(128, 128, 128, 2, dropout=0.0, Normalization='ln', InputNorm=True)
    x = torch.randn(9000, 128)
    with torch.no_grad():
        assert not m._tc_ok(x)                       # tc_dtype unset
        m.tc_dtype = torch.bfloat16
        assert not m._tc_ok(x)                       # CPU rows
    conv = allset_b200.HalfNLHconv(128, 128, 128, 2, 0.0, 'ln', True, heads=8, attention=True)
    conv.set_agg_dtype(torch.bfloat16)
    assert conv.prop.rFF.tc_dtype == torch.bfloat16 and conv.prop.agg_dtype == torch.bfloat16
    with torch.no_grad():
        assert not conv.prop._tc_v_ok(x)
    conv.set_agg_dtype(None)
    assert conv.prop.rFF.tc_dtype is None
    ds = allset_b200.HalfNLHconv(128, 128, 128, 2, 0.0, 'ln', True, heads=1, attention=False)
    ds.set_agg_dtype(torch.bfloat16)
    assert ds.f_enc.tc_dtype == torch.bfloat16 and ds.f_dec.tc_dtype == torch.bfloat16
    odd = allset_b200.MLP(128, 64, 128, 2, dropout=0.0, Normalization='ln', InputNorm=True)
    odd.tc_dtype = torch.bfloat16
    with torch.no_grad():
        assert not odd._tc_ok(x)                     # not square


def test_tcgen05_linear_gating_is_cuda_square_only():
    """ops.linear_nb / MLP._forward_chain hand a Linear to the tcgen05 kernels only for CUDA rows, square widths 64 / 128,
    f32 master weights and at least FUSED_DENSE_MIN_ROWS rows (the module API itself refuses CPU tensors: test_no_cpu_fallback)."""
    from allset_b200 import _lib, ops
    x = torch.randn(9000, 128)
    w = torch.randn(128, 128)
    assert not _lib.linear_ok(x, w) and not ops.tc_linear_ok(x, w)          # CPU rows
    assert _lib.LINEAR_WIDTHS == (64, 128) and (_lib.PREC_BF16, _lib.PREC_SPLIT) == (0, 1)
    hdr = open(os.path.join(ROOT, 'include', 'allset_b200.h')).read()
    assert '#define ALLSET_PREC_BF16 0' in hdr and '#define ALLSET_PREC_SPLIT 1' in hdr


# ----------------------------------------------------------------------------------------------------------------------
# Work split of the stream kernels (csrc/allset_kernels.cu: chunk_boundary / chunk_cut), restated in Python: the
# invariants the decoupled look-back relies on.  A model check of the ALGORITHM (the device code is exercised on the GPU).
# ----------------------------------------------------------------------------------------------------------------------
def _cuts(rowptr, n_chunks, split_min_len=256, split_min_piece=32):
    import bisect
    n = len(rowptr) - 1
    cost = [rowptr[s] + s for s in range(n + 1)]
    total = cost[n]
    out = []
    for c in range(n_chunks + 1):
        if c <= 0:
            out.append((0, rowptr[0], 0)); continue
        if c >= n_chunks:
            out.append((n, rowptr[n], 0)); continue
        target = total * c // n_chunks
        hi = 0 if target <= 0 else bisect.bisect_left(cost, target)          # smallest s with cost(s) >= target
        seg, k, inside = hi, rowptr[hi], 0
        if hi > 0:
            s = hi - 1
            beg, ln = rowptr[s], rowptr[hi] - rowptr[s]
            o = target - (beg + s)
            if ln >= split_min_len:
                if o < split_min_piece:
                    seg, k = s, beg
                elif ln - o >= split_min_piece:
                    seg, k, inside = s, beg + o, 1
        out.append((seg, k, inside))
    return out


@pytest.mark.parametrize('seed', range(6))
def test_stream_chunk_cuts_cover_every_incidence_once_and_chain_consistently(seed):
    import random
    rnd = random.Random(seed)
    n = rnd.choice([400, 2000, 5000])
    lens = [rnd.choice([0, 1, 2, 3, 5, 8]) for _ in range(n)]
    for _ in range(rnd.choice([3, 20, 80])):                                    # long segments, some adjacent
        i = rnd.randrange(n - 2)
        lens[i] = rnd.choice([256, 300, 1000, 4096, 20000])
        if rnd.random() < 0.5:
            lens[i + 1] = rnd.choice([257, 4096])
    rowptr = [0]
    for ln in lens:
        rowptr.append(rowptr[-1] + ln)
    n_chunks = max(2, n // 16)
    cuts = _cuts(rowptr, n_chunks)
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert a[0] <= b[0] and a[1] <= b[1], (a, b)                             # monotone in segment and position
        if b[2]:
            assert rowptr[b[0]] + 32 <= b[1] <= rowptr[b[0] + 1] - 32            # both pieces keep >= 32 incidences
    covered = 0
    published = {}                                                               # chunk -> flag
    for c in range(n_chunks):
        (s0, k0, in0), (s1, k1, in1) = cuts[c], cuts[c + 1]
        covered += k1 - k0
        nseg = s1 - s0
        assert nseg >= 0
        if in1:
            published[c] = 2 if (in0 and nseg == 0) else 1
            assert k1 > k0                                                       # a publisher never returns early
        if in0 and nseg > 0:                                                     # owner: walk back to the first piece
            q = c - 1
            while published[q] != 1:                                             # KeyError = waiting for a piece nobody publishes
                assert cuts[q][0] == s0 and cuts[q][2]                           # middle chunks lie inside the same segment
                q -= 1
            assert cuts[q + 1][0] == s0
    assert covered == rowptr[-1]
