"""Kernel-level profile of one SetGNN training step (forward + backward + Adam) at config-3 size, fp32 and bf16 mode.
python scripts/prof_train.py [rows_limit]  -> torch.profiler tables (CUDA time by kernel) + ms per step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from types import SimpleNamespace
import torch, allset_b200, allset_oracle as O
from allset_b200 import synthetic, preprocessing as P
from torch.profiler import profile, ProfilerActivity

n, m, d = 1_000_000, 200_000, 128
v2e = synthetic.poisson_hypergraph(n, m, 20, seed=1234, device='cuda:0')
ei, tot = P.add_self_loops(v2e, n, m); norm = P.norm_construction(ei)
x = synthetic.features(n, d, torch.float32, device='cuda:0')
y = torch.randint(0, 10, (n,), device='cuda:0')
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 24
only = sys.argv[2] if len(sys.argv) > 2 else ''          # 'f32only' | 'bf16only' | 'deepsets' (short runs under ncu)
for pma, heads in ((False, 1), (True, 8)):
    args = O.config_namespace(num_features=d, num_classes=10, MLP_hidden=d, Classifier_hidden=d, heads=heads,
                              All_num_layers=1, Classifier_num_layers=1, PMA=pma, aggregate='add')
    if only == 'deepsets' and pma:
        continue
    for agg in (None, torch.bfloat16):
        if (only == 'f32only' and agg is not None) or (only == 'bf16only' and agg is None):
            continue
        torch.manual_seed(0)
        model = allset_b200.SetGNN(args, agg_dtype=agg).to('cuda:0').train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        data = SimpleNamespace(x=x, edge_index=ei.clone(), norm=norm)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.nll_loss(torch.log_softmax(model(data), dim=1), y)
            loss.backward(); opt.step()

        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): step()
        e1.record(); torch.cuda.synchronize()
        print('==== %s agg_dtype=%s: %.3f ms per training step' % ('AllSetTransformer' if pma else 'AllDeepSets', agg,
                                                                  e0.elapsed_time(e1) / 10), flush=True)
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3): step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=limit, max_name_column_width=80))
        del model, opt
        torch.cuda.empty_cache()
