"""Kernel micro-benchmark for tuning (not the contract bench): times V->E / E->V of the sum and PMA kernels on one
synthetic graph for the library given by ALLSET_B200_LIB.  python scripts/kbench.py [nodes hyperedges mean d]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import allset_b200
from allset_b200 import _lib, synthetic, sharding

def main():
    a = sys.argv[1:]
    Nv, Me, mean, d = (int(a[0]), int(a[1]), float(a[2]), int(a[3])) if len(a) >= 4 else (10_000_000, 2_000_000, 30.0, 128)
    dt = torch.bfloat16 if os.environ.get('KB_DTYPE', 'bf16') == 'bf16' else torch.float32
    es = 2 if dt == torch.bfloat16 else 4
    dev = torch.device('cuda:0')
    torch.zeros(1, device=dev)
    if os.environ.get('KB_L2_FETCH'):
        import ctypes, glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), 'lib', 'libcudart*.so*')) + glob.glob('/usr/local/cuda/lib64/libcudart.so*')
        rt = ctypes.CDLL(cands[0])
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ['KB_L2_FETCH'])))
        val2 = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val2), 5)
        print('L2 fetch granularity: was', val.value, 'set rc', rc, 'now', val2.value, 'via', cands[0], flush=True)
    if os.environ.get('KB_GRAPH') == 'powerlaw':
        ei = synthetic.powerlaw_hypergraph(Nv, Me, 2, 4096, 2.0, seed=1234, device=dev)
    else:
        ei = synthetic.poisson_hypergraph(Nv, Me, mean, seed=1234, device=dev)
    v2e = allset_b200.Incidence.from_coo(ei[0], ei[1] - Nv, n_src=Nv, n_tgt=Me)
    del ei
    sh = sharding.ShardedIncidence(v2e, 0, 1)
    x_v = synthetic.features(Nv, d, dt, device=dev); x_e = torch.empty(Me, d, dtype=dt, device=dev); x_v2 = torch.empty_like(x_v)
    H = int(os.environ.get('KB_HEADS', '8'))
    sv = torch.randn(Nv, H, device=dev); se = torch.randn(Me, H, device=dev); seed = torch.randn(d, device=dev)
    def timeit(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    nnz = v2e.nnz
    res = {'lib': os.path.basename(_lib.LIB_PATH), 'graph': os.environ.get('KB_GRAPH', 'poisson'), 'N': Nv, 'M': Me, 'd': d, 'dtype': str(dt), 'nnz': nnz,
           'long_segments_v2e': 0 if v2e.by_tgt.long_ids is None else int(v2e.by_tgt.long_ids.numel())}
    for name, fn, b in (
        ('sum_v2e', lambda: sh.v2e_reduce(x_v, x_e), synthetic.algorithmic_bytes(nnz, Me, d, es)),
        ('sum_e2v', lambda: sh.e2v_reduce(x_e, x_v2), synthetic.algorithmic_bytes(nnz, Nv, d, es)),
        ('mean_e2v', lambda: sh.e2v_reduce(x_e, x_v2, True), synthetic.algorithmic_bytes(nnz, Nv, d, es)),
        ('pma_v2e', lambda: sh.v2e_pma(x_v, sv, seed, H, x_e), synthetic.algorithmic_bytes(nnz, Me, d, es, heads=H)),
        ('pma_e2v', lambda: sh.e2v_pma(x_e, se, seed, H, x_v2), synthetic.algorithmic_bytes(nnz, Nv, d, es, heads=H)),
    ):
        if os.environ.get('KB_ONLY') and name not in os.environ['KB_ONLY'].split(','):
            continue
        ms = timeit(fn, int(os.environ.get('KB_ITERS', '20')))
        res[name] = {'ms': round(ms, 4), 'gbs': round(b / ms / 1e6, 1), 'frac': round(b / ms / 1e6 / 6464.9, 4)}
    print(json.dumps(res), flush=True)

main()
