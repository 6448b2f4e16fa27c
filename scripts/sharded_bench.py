"""SetGNN over a partitioned hypergraph: forward (eval) and training-step time at N ranks, as named by BASELINE.json
configs[3] (|V|=10M |E|=2M mean-deg 30, AllSetTransformer d=128 heads=8, 4 x B200) and configs[4] (power-law |V|=50M |E|=8M,
AllDeepSets d=256 bf16, 8 x B200).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/sharded_bench.py --config cfg4|cfg5|small [--layers L] [--train]

Prints one JSON line on rank 0: ms per forward / per training step (CUDA events, max over ranks), hyperedges/s, and a
check of the sharded logits of the owned rows against an unsharded forward on the same graph when that fits (`--verify`).
The graph has NO self-loop hyperedges added (the raw synthetic incidence, as in bench.py), so |E| is the configured one.
"""
import argparse
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

CONFIGS = {
    'cfg4': dict(nodes=10_000_000, hyperedges=2_000_000, graph='poisson', mean=30.0, d=128, pma=True, heads=8),
    'cfg5': dict(nodes=50_000_000, hyperedges=8_000_000, graph='powerlaw', mean=0.0, d=256, pma=False, heads=1),
    'cfg5s': dict(nodes=20_000_000, hyperedges=3_200_000, graph='powerlaw', mean=0.0, d=256, pma=False, heads=1),
    'small': dict(nodes=1_000_000, hyperedges=200_000, graph='poisson', mean=20.0, d=128, pma=True, heads=8),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='small', choices=sorted(CONFIGS))
    ap.add_argument('--layers', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--train', action='store_true')
    ap.add_argument('--verify', action='store_true')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--full-xv', action='store_true', help='send every vertex row to every rank (no per-row peer mask)')
    a = ap.parse_args()
    c = CONFIGS[a.config]

    import torch
    import torch.distributed as dist
    import allset_b200
    import allset_oracle as O                          # only for the args namespace helper
    from allset_b200 import synthetic
    from allset_b200.sharded_model import ShardedSetGNN

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    Nv, Me, d = c['nodes'], c['hyperedges'], c['d']
    if c['graph'] == 'powerlaw':
        ei = synthetic.powerlaw_hypergraph(Nv, Me, 2, 4096, 2.0, seed=1234, device=dev)
    else:
        ei = synthetic.poisson_hypergraph(Nv, Me, c['mean'], seed=1234, device=dev)
    nnz = int(ei.shape[1])
    agg = torch.bfloat16 if a.dtype == 'bf16' else None
    args = O.config_namespace(num_features=d, num_classes=10, MLP_hidden=d, Classifier_hidden=d, heads=c['heads'],
                              All_num_layers=a.layers, Classifier_num_layers=1, PMA=c['pma'], aggregate='add', dropout=0.5)
    torch.manual_seed(0)
    model = allset_b200.SetGNN(args, agg_dtype=agg).to(dev)
    sm = ShardedSetGNN(model, selective=not a.full_xv)
    # every rank materialises only the feature rows it owns: ShardedSetGNN takes `x` as the owned rows when data.num_nodes
    # is given (the full [N, F] fp32 matrix of config 5 would be 51 GB per rank for nothing)
    data = SimpleNamespace(x=None, edge_index=ei, norm=torch.ones(nnz, dtype=torch.int64, device=dev), num_nodes=Nv)
    sh, _, _ = sm._directions(SimpleNamespace(x=torch.empty(Nv, 0, device=dev), edge_index=ei, norm=None))
    data.x = synthetic.features(sh.v_hi - sh.v_lo, d, torch.float32, seed=1234 + rank, device=dev)
    y = torch.randint(0, 10, (sh.v_hi - sh.v_lo,), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    model.eval()

    def fwd():
        with torch.no_grad():
            return sm(data)

    fwd_ms = timed(fwd, a.steps)
    res = {'config': a.config, 'n_gpus': world, 'nodes': Nv, 'hyperedges': Me, 'nnz': nnz, 'd': d, 'layers': a.layers,
           'model': 'AllSetTransformer heads=%d' % c['heads'] if c['pma'] else 'AllDeepSets', 'dtype': a.dtype,
           'xv_exchange': 'selective (per-row peer mask)' if not a.full_xv else 'every row to every rank',
           'fwd_ms': fwd_ms, 'fwd_hyperedges_per_s': Me * a.layers / (fwd_ms * 1e-3)}
    if a.train:
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.nll_loss(torch.log_softmax(sm(data), dim=1), y)
            loss.backward()
            sm.allreduce_gradients()
            opt.step()

        res['train_step_ms'] = timed(step, max(3, a.steps // 2))
    if a.verify and world > 1:
        # unsharded forward of the same model on the same graph needs the full feature matrix: small configs only
        model.eval()
        xs = [torch.empty_like(data.x) if (r_hi - r_lo) == data.x.shape[0] else torch.empty((r_hi - r_lo, d), device=dev)
              for r_lo, r_hi in sh.v_ranges]
        dist.all_gather(xs, data.x) if len({t.shape for t in xs}) == 1 else None
        if len({t.shape for t in xs}) == 1:
            full_x = torch.cat(xs)
            with torch.no_grad():
                ref = model(SimpleNamespace(x=full_x, edge_index=ei, norm=data.norm))
                mine = sm(data)
            err = (ref[sh.v_lo:sh.v_hi] - mine).abs().max()
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            res['max_abs_err_vs_unsharded'] = float(err)
            res['logit_scale'] = float(ref.abs().max())
    res['mem_gb'] = torch.cuda.max_memory_allocated() / 1e9
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
