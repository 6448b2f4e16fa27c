"""bf16-mode error budget on the reference-generated goldens (cora AllDeepSets, citeseer AllSetTransformer): the error
of the GPU path against the fp32 reference next to the INHERENT error of the mode -- the reference's own arithmetic
with only the gathered rows and the aggregation outputs rounded to bf16 (CPU oracle) -- per half-layer tap and at the
logits.  Test tooling (imports oracle/): python scripts/bf16_parity_report.py -> JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import allset_oracle as O
import allset_b200
from allset_b200 import ops
from conftest import load_golden, golden_x
from test_gpu_parity import _build, _taps

bf = lambda t: t.to(torch.bfloat16).float()
orig_sum, orig_pma = O.aggregate_sum_mean, O.aggregate_pma


def inherent(rec):
    a = rec['args']
    O.aggregate_sum_mean = lambda x, s, t, n, g: bf(orig_sum(bf(x), s, t, n, g))
    O.aggregate_pma = lambda v, sc, sd, s, t, sl=0.2: (lambda o: (bf(o[0]), o[1]))(orig_pma(bf(v), sc, sd, s, t, sl))
    try:
        with torch.no_grad():
            return O.setgnn(rec['state_dict'], golden_x(rec), rec['edge_index'], rec['norm'], PMA=a['PMA'], heads=a['heads'],
                            aggregate=a['aggregate'])
    finally:
        O.aggregate_sum_mean, O.aggregate_pma = orig_sum, orig_pma


for name in ('cora_alldeepsets.pt', 'citeseer_allsettransformer.pt'):
    rec = load_golden(name)
    s = rec['tap_stride']
    q_out, q_taps = inherent(rec)
    scale = rec['logits'].abs().max().item()
    for mode, min_rows in (('bf16 storage, fp32 dense (small graph: ATen dense)', None), ('bf16 mode, tcgen05 / chain dense', 0)):
        if min_rows is not None:
            ops.FUSED_DENSE_MIN_ROWS = min_rows
        model, data = _build(rec, agg_dtype=torch.bfloat16)
        taps, hooks = _taps(model)
        with torch.no_grad():
            out = model(data)
        for h in hooks:
            h.remove()
        line = {'golden': name, 'mode': mode, 'logit_scale': scale,
                'gpu_vs_ref': (out.cpu() - rec['logits']).abs().max().item() / scale,
                'inherent_vs_ref': (q_out - rec['logits']).abs().max().item() / scale,
                'gpu_vs_inherent': (out.cpu() - q_out).abs().max().item() / scale, 'taps': []}
        for mine, ref, q in zip(taps, rec['taps'], q_taps):
            r = torch.relu(ref)
            ts = r.abs().max().item()
            line['taps'].append({'gpu_vs_ref': (mine[::s] - r).abs().max().item() / ts,
                                 'inherent_vs_ref': (q[::s] - r).abs().max().item() / ts})
        print(json.dumps(line), flush=True)
