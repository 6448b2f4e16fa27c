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
for pma, heads in ((False, 1), (True, 8)):
    args = O.config_namespace(num_features=d, num_classes=10, MLP_hidden=d, Classifier_hidden=d, heads=heads, All_num_layers=1, Classifier_num_layers=1, PMA=pma, aggregate='add')
    torch.manual_seed(0); model = allset_b200.SetGNN(args).to('cuda:0').train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    data = SimpleNamespace(x=x, edge_index=ei.clone(), norm=norm)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(model(data), dim=1), y)
        loss.backward(); opt.step()
    for _ in range(3): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    print('PMA' if pma else 'DeepSets')
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=90))
