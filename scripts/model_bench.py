"""SetGNN forward / forward+backward latency per BASELINE.json config, B200 module API vs the CPU oracle.
python scripts/model_bench.py [--cpu] -> JSON lines.  Real-data configs come from tests/golden (recorded reference
state_dicts); the synthetic config uses reset_parameters() under torch.manual_seed(0)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from types import SimpleNamespace
import torch
import allset_b200
import allset_oracle as O
from allset_b200 import synthetic


def densify(sp):
    x = torch.zeros(sp['shape'], dtype=sp['vals'].dtype)
    x[sp['rows'].long(), sp['cols'].long()] = sp['vals']
    return x


def time_gpu(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_cpu(fn, iters=5, warm=1):
    for _ in range(warm): fn()
    t = time.perf_counter()
    for _ in range(iters): fn()
    return (time.perf_counter() - t) / iters * 1e3


def run(name, args, x, ei, norm, state_dict, agg_dtype, n_he, do_cpu):
    dev = torch.device('cuda:0')
    model = allset_b200.SetGNN(args, agg_dtype=agg_dtype)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    else:
        torch.manual_seed(0); model.reset_parameters()
    model.to(dev)
    data = SimpleNamespace(x=x.to(dev), edge_index=ei.clone().to(dev), norm=norm.to(dev))
    model.eval()
    with torch.no_grad():
        model(data)
        fwd = time_gpu(lambda: model(data))
    graphed = None
    try:
        g = allset_b200.GraphedForward(model, data)
        graphed = time_gpu(lambda: g())
    except Exception as e:  # noqa
        graphed = repr(e)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    y = torch.randint(0, args.num_classes, (x.shape[0],), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(model(data), dim=1), y)
        loss.backward(); opt.step()
    train = time_gpu(step, iters=20, warm=3)
    rec = {'config': name, 'agg_dtype': str(agg_dtype), 'fwd_ms': fwd, 'fwd_cuda_graph_ms': graphed, 'train_step_ms': train,
           'hyperedges': n_he, 'fwd_hyperedges_per_s': n_he / (fwd * 1e-3)}
    if do_cpu:
        params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        xc, eic, nc = x.float(), ei, norm
        with torch.no_grad():
            cpu = time_cpu(lambda: O.setgnn(params, xc, eic, nc, PMA=args.PMA, heads=args.heads, aggregate=args.aggregate))
        rec.update({'cpu_oracle_fwd_ms': cpu, 'cpu_cores': os.cpu_count(), 'fwd_speedup_vs_cpu_oracle': cpu / fwd})
    print(json.dumps(rec), flush=True)


def main():
    do_cpu = '--cpu' in sys.argv
    for fname, nm in (('cora_alldeepsets.pt', 'configs[0] cora AllDeepSets d=64 L=1'),
                      ('citeseer_allsettransformer.pt', 'configs[1] citeseer AllSetTransformer d=128 heads=4 L=2')):
        rec = torch.load(os.path.join(ROOT, 'tests', 'golden', fname), weights_only=True)
        args = SimpleNamespace(**rec['args'])
        n_he = int(rec['edge_index'][1].max() - rec['edge_index'][1].min()) + 1
        for dt in (None, torch.bfloat16):
            run(nm, args, densify(rec['x_sparse']), rec['edge_index'], rec['norm'], rec['state_dict'], dt, n_he, do_cpu and dt is None)
    # configs[2]: synthetic 1M / 200K, AllDeepSets d=128, reference pipeline incl. self-loops
    n, m, d = 1_000_000, 200_000, 128
    from allset_b200 import preprocessing as P
    v2e = synthetic.poisson_hypergraph(n, m, 20, seed=1234, device='cuda:0')
    ei, tot = P.add_self_loops(v2e, n, m)
    norm = P.norm_construction(ei)
    x = synthetic.features(n, d, torch.float32, device='cuda:0')
    for pma, heads, nm in ((False, 1, 'configs[2] synthetic 1M/200K(+1M self-loops) AllDeepSets d=128 L=1'),
                           (True, 8, 'synthetic 1M/200K(+1M self-loops) AllSetTransformer d=128 heads=8 L=1')):
        args = O.config_namespace(num_features=d, num_classes=10, MLP_hidden=d, Classifier_hidden=d, heads=heads,
                                  All_num_layers=1, Classifier_num_layers=1, PMA=pma, aggregate='add')
        for dt in (None, torch.bfloat16):
            run(nm, args, x, ei, norm, None, dt, tot, do_cpu and dt is None)

def uni_bench():
    """UniGCNIIConv (reference src/models.py:909-942) at config-3 size: the two segmented reduces vs the ATen
    index_add_ (atomic scatter) chain the reference's torch_scatter path amounts to on a GPU."""
    dev = torch.device('cuda:0')
    n, m, d = 1_000_000, 200_000, 128
    ei = synthetic.poisson_hypergraph(n, m, 20, seed=1234, device=dev)
    V, E = ei[0].contiguous(), (ei[1] - n).contiguous()
    degV = torch.bincount(V, minlength=n).view(-1, 1).float()
    cnt = torch.bincount(E, minlength=m).view(-1, 1).float()
    degE = (torch.zeros(m, 1, device=dev).index_add_(0, E, degV[V]) / cnt.clamp(min=1)).pow(-0.5)
    degV = degV.pow(-0.5); degV[torch.isinf(degV)] = 1
    args = SimpleNamespace(UniGNN_degV=degV, UniGNN_degE=degE, UniGNN_use_norm=False)
    conv = allset_b200.UniGCNIIConv(args, d, d).to(dev)
    x = torch.randn(n, d, device=dev)

    def atomics():
        xe = torch.zeros(m, d, device=dev).index_add_(0, E, x[V]) / cnt.clamp(min=1) * degE
        xv = torch.zeros(n, d, device=dev).index_add_(0, V, xe[E]) * degV
        xi = 0.9 * xv + 0.1 * x
        return 0.7 * xi + 0.3 * conv.W(xi)
    with torch.no_grad():
        ours = time_gpu(lambda: conv(x, V, E, 0.1, 0.3, x))
        ref = time_gpu(atomics, iters=10, warm=2)
    print(json.dumps({'config': 'UniGCNIIConv 1M/200K d=128 fp32 (one layer)', 'fwd_ms': ours,
                      'aten_gather_index_add_chain_ms': ref, 'speedup': ref / ours}), flush=True)


main()
uni_bench()
