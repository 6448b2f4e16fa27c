"""Per-kernel timing of the tcgen05 Linear kernels against the cuBLAS calls they replace (1 M rows, d = 128 / 64):
forward, input gradient, weight gradient; fp32 rows (split precision vs SIMT SGEMM) and bf16 rows.  JSON lines with
ms, algorithmic GB/s (rows read once + rows written once) and the fraction of the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from allset_b200 import _lib

peak = 6547.5
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', peak))
except Exception:
    pass


def t(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


dev = torch.device('cuda:0')
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for d in (128, 64):
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(rows, d, device=dev).to(dt)
        dy = torch.randn(rows, d, device=dev).to(dt)
        w = torch.randn(d, d, device=dev) / d ** 0.5
        b = torch.randn(d, device=dev)
        wc = w.to(dt)
        s = x.element_size()
        io = 2 * rows * d * s
        rec = {'rows': rows, 'd': d, 'dtype': str(dt).replace('torch.', ''), 'hbm_peak_gbs': peak}
        for name, ours, lib, nbytes in (
                ('fwd', lambda: _lib.linear_fwd(x, w), lambda: torch.mm(x, wc.t()), io),
                ('fwd_bias_relu', lambda: _lib.linear_fwd(x, w, b, relu=True), lambda: torch.relu_(torch.addmm(b.to(dt), x, wc.t())), io),
                ('dgrad', lambda: _lib.linear_fwd(dy, w, transposed=True), lambda: torch.mm(dy, wc), io),
                ('wgrad', lambda: _lib.linear_wgrad(dy, x),
                 (lambda: torch.mm(dy.t(), x)) if dt == torch.float32 else (lambda: torch.mm(dy.t(), x, out_dtype=torch.float32)), io)):
            ms, ms_lib = t(ours), t(lib)
            rec[name] = {'ms': round(ms, 4), 'gbs': round(nbytes / ms / 1e6, 1), 'frac': round(nbytes / ms / 1e6 / peak, 3),
                         'cublas_ms': round(ms_lib, 4)}
        print(json.dumps(rec), flush=True)
