#!/usr/bin/env python
"""bench.py -- hyperedge-aggregations/sec (V->E + E->V) at d=128 on B200, with HBM roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the named graph: the V->E segmented reduce over every hyperedge followed
by the E->V segmented reduce over every vertex (AllDeepSets, aggregate='add' as reference src/train.py:36-38 forces),
d=128, bf16 rows / fp32 accumulate.  Workload = BASELINE.json configs[3]'s graph (synthetic |V|=10M, |E|=2M,
hyperedge size 1+Poisson(29), nnz~60M, seed 1234) -- the 10M-vertex / 2M-hyperedge synthetic the north_star target
is quoted on; it fits one B200.  With N > 1 the SAME graph is hyperedge-sharded (V->E) / vertex-sharded (E->V) over
N ranks (strong scaling).  The exchange of X_e between the two directions is inside every step.  The updated X_v is
left vertex-sharded by default (last layer: the next operator is row-parallel); `--replicate-xv` adds north_star's
per-layer replication of X_v to every step, and BOTH modes are timed in every multi-GPU run (`other_mode`).

JSON line (rank 0): value = |E| * K / t (device-timed, inputs resident in HBM, max over ranks); e2e = same metric
through the public API with the vertex features starting in pinned HOST memory and the result read back to the host
every step; roofline = the segmented-reduce kernel's algorithmic bytes / its CUDA-event duration against
MEASURED_PEAKS.json (`traffic` is the DRAM byte count of the committed ncu capture of the same launch, read from
profiles/traffic.json -- static, labelled as such); cpu_baseline = the reference's CPU op sequence (oracle port:
index_select -> norm*x_j -> scatter_add_, fp32) on a 1/10-scale graph of the same distribution on this box's host
cores, with the full half-layer pair (f_enc -> aggregate -> f_dec, twice) timed beside it.

--impl reference times that CPU path alone (the reference has no GPU kernel of its own).
Only the cpu_baseline / --impl reference legs import oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'hyperedge-aggregations/sec (V->E+E->V) at d=128'
UNIT = 'hyperedges/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--nodes', type=int, default=10_000_000, help='|V|')
    ap.add_argument('--hyperedges', type=int, default=2_000_000, help='|E|')
    ap.add_argument('--mean-size', type=float, default=30.0)
    ap.add_argument('--graph', default='poisson', choices=['poisson', 'powerlaw'],
                    help='hyperedge sizes: 1+Poisson(mean-1) (configs 3, 4) or P(s)~s^-2 on [2,4096] with a forced 4096 (config 5)')
    ap.add_argument('--d', '--width', dest='d', type=int, default=128,
                    help='feature width (use --width under torchrun: its own parser treats --d as an ambiguous prefix)')
    ap.add_argument('--heads', type=int, default=8)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--cpu-scale', type=int, default=10, help='CPU legs run on a 1/scale graph of the same distribution')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-pma', action='store_true')
    ap.add_argument('--no-mlp', action='store_true')
    ap.add_argument('--exchange', default='auto', choices=['auto', 'fused', 'nccl'],
                    help='N>1: fused = P2P stores from the kernel epilogue into symmetric memory; nccl = all-gather after')
    ap.add_argument('--full-xv', action='store_true',
                    help='N>1 with X_v replication: send every updated vertex row to every rank (north_star\'s plain all-gather) '
                         'instead of only to the ranks that gather it')
    ap.add_argument('--replicate-xv', action='store_true',
                    help='N>1: also replicate the updated X_v inside every step (needed only when another layer follows)')
    return ap.parse_args()


def workload_name(a):
    if a.graph == 'powerlaw':
        return ('synthetic power-law |V|=%s |E|=%s max-deg=4096 AllDeepSets(sum) V->E+E->V d=%d %s'
                % (_si(a.nodes), _si(a.hyperedges), a.d, a.dtype))
    return ('synthetic |V|=%s |E|=%s mean-deg=%g AllDeepSets(sum) V->E+E->V d=%d %s'
            % (_si(a.nodes), _si(a.hyperedges), a.mean_size, a.d, a.dtype))


def make_graph(a, n, m, device):
    from allset_b200 import synthetic
    if a.graph == 'powerlaw':
        return synthetic.powerlaw_hypergraph(n, m, 2, 4096, 2.0, seed=a.seed, device=device)
    return synthetic.poisson_hypergraph(n, m, a.mean_size, seed=a.seed, device=device)


def _si(n):
    for div, suf in ((1_000_000, 'M'), (1_000, 'K')):
        if n >= div and n % (div // 10) == 0:
            v = n / div
            return ('%d%s' % (v, suf)) if v == int(v) else ('%g%s' % (v, suf))
    return str(n)


# ------------------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------------------------
_REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
            0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}


class ClockSampler(threading.Thread):
    """Samples SM clock + clock-event reasons of one GPU through NVML every `period` seconds while running."""

    def __init__(self, device_index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.sm_max = period, [], set(), None
        self._halt = threading.Event()
        self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(device_index).uuid)
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except Exception:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa
            self.error = repr(e)
            self.handle = None

    def run(self):
        if self.handle is None:
            return
        nv = self.nv
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                bits = int(get_reasons(self.handle))
                for bit, name in _REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        out = {'sm_mhz': (statistics.median(self.samples) if self.samples else None), 'sm_max_mhz': self.sm_max,
               'reasons': sorted(self.reasons), 'samples': len(self.samples)}
        if self.handle is None:
            out['error'] = getattr(self, 'error', 'nvml unavailable')
        return out


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (the only code in this file that touches oracle/)
# ------------------------------------------------------------------------------------------------------------------
def cpu_sample_graph(a):
    import torch
    from allset_b200 import synthetic
    n, m = max(a.nodes // a.cpu_scale, 1000), max(a.hyperedges // a.cpu_scale, 200)
    ei = make_graph(a, n, m, 'cpu')
    node, he = ei[0], ei[1] - n
    x = torch.randn(n, a.d, generator=torch.Generator().manual_seed(a.seed))
    norm = torch.ones(ei.shape[1], dtype=torch.int64)        # data.norm = ones_like(edge_index[0]) (int64), preprocessing.py:454
    desc = ('1/%d-scale graph of the same distribution (|V|=%d |E|=%d nnz=%d, d=%d), fp32, reference op sequence '
            'index_select -> norm*x_j -> scatter_add_ per direction' % (a.cpu_scale, n, m, ei.shape[1], a.d))
    return n, m, node, he, x, norm, desc


def cpu_time_pairs(a, steps, warmup, budget_s=None):
    """Times `steps` V->E + E->V pairs of the oracle's reference op sequence; returns (hyperedges/s, cores, desc, ms/step, steps)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import allset_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, m, node, he, x, norm, desc = cpu_sample_graph(a)
    for _ in range(warmup):
        O.layer_pair_sum(x, node, he, norm, 'sum')
    done, t0 = 0, time.perf_counter()
    for _ in range(steps):
        O.layer_pair_sum(x, node, he, norm, 'sum')
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return m * done / dt, cores, desc, dt / done * 1e3, done


def cpu_time_half_layers(a, steps=2, budget_s=15.0):
    """The full AllDeepSets half-layer pair of the reference on the same sample graph: relu(f_enc) -> aggregate ->
    relu(f_dec), V->E then E->V (oracle.half_nlh_conv = reference src/layers.py:623-636), random-init weights, eval."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import allset_oracle as O
    n, m, node, he, x, norm, _ = cpu_sample_graph(a)
    g = torch.Generator().manual_seed(a.seed + 7)
    d = a.d
    params = {}
    for conv in ('V2EConvs.0.', 'E2VConvs.0.'):
        for f in ('f_enc.', 'f_dec.'):
            for i in range(2):
                params[conv + f + 'lins.%d.weight' % i] = torch.randn(d, d, generator=g) / d ** 0.5
                params[conv + f + 'lins.%d.bias' % i] = torch.zeros(d)
                params[conv + f + 'normalizations.%d.weight' % i] = torch.ones(d)
                params[conv + f + 'normalizations.%d.bias' % i] = torch.zeros(d)

    def pair():
        with torch.no_grad():
            xe = O.half_nlh_conv(params, 'V2EConvs.0.', x, node, he, norm, 'sum', attention=False)
            return O.half_nlh_conv(params, 'E2VConvs.0.', xe, he, node, norm, 'sum', attention=False)

    pair()
    done, t0 = 0, time.perf_counter()
    for _ in range(steps):
        pair()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return {'value': m / dt, 'unit': UNIT, 'ms_per_pair_on_sample': dt * 1e3, 'steps': done,
            'what': 'reference half layers (LN->Linear->ReLU->LN->Linear, relu, aggregate, same again) x2, fp32, eval'}


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    value, cores, desc, ms, done = cpu_time_pairs(a, a.steps, a.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': done,
        'warmup': a.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(a), 'sample': desc, 'l2': 'inputs larger than L2'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def bind_host_memory_near_gpu(dev_index):
    """Pinned staging buffers should live on the NUMA node the GPU hangs off: with 8 ranks allocating on one node every
    H2D / D2H crosses the socket interconnect (round 1: the 8-GPU e2e moved 60 GB/s in aggregate).  Best effort, no
    dependency: reads the GPU's node from sysfs and sets this thread's memory policy to PREFER it (raw set_mempolicy
    syscall; libnuma is not in the image).  Returns what it did for the JSON line."""
    import ctypes
    import torch
    info = {'gpu_numa_node': None, 'nodes': None, 'policy': 'unchanged'}
    try:
        pr = torch.cuda.get_device_properties(dev_index)
        bus = '%04x:%02x:%02x.0' % (getattr(pr, 'pci_domain_id', 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        nodes = [n for n in os.listdir('/sys/devices/system/node') if n.startswith('node')]
        info.update(gpu_numa_node=node, nodes=len(nodes), pci=bus)
        if node >= 0 and len(nodes) > 1:
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))      # set_mempolicy(MPOL_PREFERRED, ...)
            info['policy'] = 'MPOL_PREFERRED node %d' % node if rc == 0 else 'set_mempolicy failed errno %d' % ctypes.get_errno()
            try:
                cpus = set()
                for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
                    lo, _, hi = part.partition('-')
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                allowed = cpus & os.sched_getaffinity(0)
                if allowed:
                    os.sched_setaffinity(0, allowed)
                    info['cpus'] = '%d cpus of node %d' % (len(allowed), node)
            except Exception as exc:  # noqa
                info['cpus'] = repr(exc)
    except Exception as exc:  # noqa
        info['error'] = repr(exc)
    return info


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import allset_b200
    from allset_b200 import _lib, sharding, synthetic

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py --impl b200 needs a CUDA device (allset_b200 has no CPU path)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner on STDOUT at communicator creation; the contract is ONE JSON line there
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    if world != a.gpus and rank == 0:
        sys.stderr.write('warning: --gpus %d but WORLD_SIZE %d; using WORLD_SIZE\n' % (a.gpus, world))
    _lib.lib()                                            # fail loudly here if the CUDA library is missing

    dtype = torch.bfloat16 if a.dtype == 'bf16' else torch.float32
    es = 2 if a.dtype == 'bf16' else 4
    Nv, Me, d, H = a.nodes, a.hyperedges, a.d, a.heads

    # ---- graph (resident; built once like the reference's data.to(device)) ------------------------------------
    ei = make_graph(a, Nv, Me, dev)
    he = ei[1] - Nv
    v2e = allset_b200.Incidence.from_coo(ei[0], he, n_src=Nv, n_tgt=Me)
    nnz = v2e.nnz
    del ei, he
    sh = sharding.ShardedIncidence(v2e, rank, world)
    torch.cuda.empty_cache()

    x_v = synthetic.features(Nv, d, dtype, seed=a.seed, device=dev)
    exchange = 'none'
    x_e = x_v2 = None
    if world > 1 and a.exchange in ('auto', 'fused'):
        try:
            x_e = sharding.ReplicatedRows(Me, d, dtype, dev)
            x_v2 = sharding.ReplicatedRows(Nv, d, dtype, dev)
            exchange = ('fused %s stores from the kernel epilogue into symmetric memory + device barrier'
                        % ('NVLS multicast' if x_e.multicast_ptr else 'P2P'))
        except Exception as e:  # noqa
            if a.exchange == 'fused':
                raise
            if rank == 0:
                sys.stderr.write('symmetric memory unavailable (%r); falling back to NCCL all-gather\n' % (e,))
            x_e = x_v2 = None
    if x_e is None:
        x_e = torch.empty((Me, d), dtype=dtype, device=dev)
        x_v2 = torch.empty((Nv, d), dtype=dtype, device=dev)
        if world > 1:
            exchange = 'NCCL all-gather after the kernel'
    plain = sharding._plain

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # ---- device-resident timing: K steps, per-phase events inside ---------------------------------------------
    def timed_steps(step_fn, n_phases, steps, warmup):
        for _ in range(warmup):
            step_fn(None)
        barrier()
        marks = [[ev() for _ in range(n_phases + 1)] for _ in range(steps)]
        start, end = ev(), ev()
        start.record()
        for k in range(steps):
            step_fn(marks[k])
        end.record()
        barrier()
        total_ms = max_over_ranks(start.elapsed_time(end))
        phases = [max_over_ranks(sum(m[i].elapsed_time(m[i + 1]) for m in marks) / steps) for i in range(n_phases)]
        return total_ms, phases

    replicate = [a.replicate_xv]

    def sum_step(marks):
        if marks: marks[0].record()
        fe = sh.v2e_reduce(x_v, x_e)
        if marks: marks[1].record()
        sh.gather_e(x_e, fe)
        if marks: marks[2].record()
        # the updated X_v goes only to the ranks whose hyperedge range gathers the row (per-row peer mask): what the
        # next layer's V->E reads, about half of north_star's full all-gather on this graph at 8 ranks
        fv = sh.e2v_reduce(x_e, x_v2 if replicate[0] else plain(x_v2), selective=selective_xv)
        if marks: marks[3].record()
        if replicate[0]:
            sh.gather_v(x_v2, fv)
        if marks: marks[4].record()

    selective_xv = world > 1 and not a.full_xv and isinstance(x_v2, sharding.ReplicatedRows)
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, ph = timed_steps(sum_step, 4, a.steps, a.warmup)
    ms_per_step = total_ms / a.steps
    value = Me / (ms_per_step * 1e-3)
    verified = None
    verify_note = None
    if world > 1:
        # every rank holds the full graph: recompute the pair unsharded and compare.  Segments shorter than the stream
        # kernels' cut threshold (256 incidences) are summed in CSR order whatever the partition, so on a graph without
        # longer ones the comparison is bit for bit.  A longer segment may be CUT at a chunk boundary, and chunk
        # boundaries depend on the slice a rank reduces: its pieces are added in a different association, the row can
        # differ by an ulp of the storage dtype, and so can everything gathered from it downstream -- then the check is
        # exact equality on the rows of short hyperedges plus a rounding-level bound on all rows.
        replicate[0] = True
        sum_step(None)
        t, sgl = v2e.by_tgt, v2e.by_src
        ref_e = _lib.segreduce_fwd(x_v, t.rowptr, t.col, t.n_tgt, False, long_ids=t.long_ids, long_threshold=t.long_threshold)
        ref_v = _lib.segreduce_fwd(ref_e, sgl.rowptr, sgl.col, sgl.n_tgt, False, long_ids=sgl.long_ids, long_threshold=sgl.long_threshold)
        len_e = (t.rowptr[1:] - t.rowptr[:-1])
        cut_possible = bool((len_e >= 256).any()) or bool(((sgl.rowptr[1:] - sgl.rowptr[:-1]) >= 256).any())
        ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -20

        def close(a_, b_):
            """|a - b| <= 4 ulp of the larger magnitude in the row (+ 4 ulp of the tensor scale for cancelling sums)"""
            a_, b_ = a_.float(), b_.float()
            bound = 4 * ulp * torch.maximum(a_.abs(), b_.abs()).amax(dim=1, keepdim=True) + 4 * ulp * b_.abs().max() * 1e-2
            return bool(((a_ - b_).abs() <= bound).all())

        def same(a_, b_, short_rows=None, what=''):
            if not cut_possible:
                return bool(torch.equal(a_, b_))
            good = close(a_, b_)
            if short_rows is not None and not bool(torch.equal(a_[short_rows], b_[short_rows])):
                good = False
            if not good:
                diff = (a_.float() - b_.float()).abs()
                rows_off = (diff.amax(dim=1) > 0)
                sys.stderr.write('[rank %d] %s: %d rows differ, max |diff| %.4g at scale %.4g%s\n' % (
                    rank, what, int(rows_off.sum()), float(diff.max()), float(b_.float().abs().max()),
                    '' if short_rows is None else ', of them short-segment rows: %d' % int((rows_off & short_rows).sum())))
            return good

        if selective_xv:
            # a rank holds its own vertex rows and the rows its hyperedge range gathers -- exactly what the next
            # layer's V->E reads: check those rows, then the next V->E itself
            need = torch.zeros(Nv, dtype=torch.bool, device=dev)
            need[sh.v_lo:sh.v_hi] = True
            need[sh.e_csr.col.long()] = True
            ok = same(plain(x_e), ref_e, len_e < 256, 'X_e') and same(plain(x_v2)[need], ref_v[need], None, 'X_v (needed rows)')
            nxt = torch.empty((sh.e_hi - sh.e_lo, d), dtype=dtype, device=dev)
            _lib.segreduce_fwd(plain(x_v2), sh.e_csr.rowptr, sh.e_csr.col, sh.e_csr.n_tgt, False, out=nxt)
            ref_nxt = _lib.segreduce_fwd(ref_v, t.rowptr, t.col, t.n_tgt, False)
            ok = ok and same(nxt, ref_nxt[sh.e_lo:sh.e_hi], None, 'next V->E')
            del nxt, ref_nxt, need
        else:
            ok = same(plain(x_e), ref_e, len_e < 256, 'X_e') and same(plain(x_v2), ref_v, None, 'X_v')
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = bool(flag.item())
        verify_note = ('bit for bit' if not cut_possible else
                       'bit for bit on the rows of hyperedges shorter than 256 incidences, within 4 ulp of the storage dtype '
                       'elsewhere (longer segments are cut at partition-dependent chunk boundaries)')
        del ref_e, ref_v, len_e
        replicate[0] = a.replicate_xv
        if not verified:
            raise RuntimeError('sharded V->E/E->V result differs from the unsharded one')
    v2e_plain_ms = None
    if world > 1:
        # the collective-free V->E segmented reduce by itself (no peer stores): the part of the path that north_star
        # expects to scale ~linearly; timed separately, not part of the step
        def v2e_only_step(marks):
            if marks: marks[0].record()
            sh.v2e_reduce(x_v, plain(x_e))
            if marks: marks[1].record()
        _, vph = timed_steps(v2e_only_step, 1, max(5, a.steps // 2), 2)
        v2e_plain_ms = vph[0]
    other = None
    if world > 1:                                  # the other exchange mode, reported beside the headline
        replicate[0] = not a.replicate_xv
        o_ms, o_ph = timed_steps(sum_step, 4, max(5, a.steps // 2), 2)
        o_steps = max(5, a.steps // 2)
        other = {'replicate_xv': replicate[0], 'xv_exchange': 'rows go to the ranks that gather them (per-row peer mask)'
                 if selective_xv else 'every row to every rank', 'value': Me / (o_ms / o_steps * 1e-3), 'unit': UNIT,
                 'ms_per_step': o_ms / o_steps,
                 'phases_ms': {'v2e': o_ph[0], 'exchange_x_e': o_ph[1], 'e2v': o_ph[2], 'exchange_x_v': o_ph[3]}}
        replicate[0] = a.replicate_xv

    # ---- end to end through the public API: host-resident features in, result read back, every step -------------
    e2e = None
    if not a.no_e2e:
        # Serving-style pipeline: every step still copies ITS input from pinned host memory and ITS result back to
        # the host, but the H2D of step k+1 and the D2H of step k-1 run on their own streams (PCIe is full duplex) while
        # step k computes.  Double-buffered on both sides; the timed region ends when the last D2H has landed.
        numa = bind_host_memory_near_gpu(local_rank) if world > 1 else None
        out_rows = sh.v_hi - sh.v_lo
        if world == 1:
            x_host = torch.empty((Nv, d), dtype=dtype).pin_memory()
            x_host.copy_(x_v)
            x_host_mine = x_host
        else:
            # a rank only ever reads ITS rows of the host matrix: pin just those (8 ranks x 2.56 GB of pinned memory on one
            # node was part of round 1's 8-GPU e2e collapse)
            x_host = None
            x_host_mine = torch.empty((out_rows, d), dtype=dtype).pin_memory()
            x_host_mine.copy_(x_v[sh.v_lo:sh.v_hi])
        out_host = [torch.empty((out_rows, d), dtype=dtype).pin_memory() for _ in range(2)]
        ph_ev = {'h2d': [], 'gather': [], 'compute': [], 'd2h': []}
        x_in_rep = None
        if world > 1 and isinstance(x_e, sharding.ReplicatedRows):
            # the input replicas live in symmetric memory: a rank lands ITS rows over PCIe and pushes them to the peers
            # itself (allset_push_rows over NVLink) instead of an NCCL all-gather competing for SMs with the reduce kernels
            x_in_rep = [sharding.ReplicatedRows(Nv, d, dtype, dev, multicast=False) for _ in range(2)]
            x_in = [r.tensor for r in x_in_rep]
        else:
            x_in = [torch.empty_like(x_v) for _ in range(2)]
        xv_out = [torch.empty_like(x_v) for _ in range(2)] if world > 1 else None
        inc_v2e, inc_e2v = v2e, v2e.reversed()
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        state = {'k': 0, 'in_done': [None, None], 'comp_done': [None, None], 'out_done': [None, None]}

        def e2e_step(_marks):
            k = state['k']; b = k % 2
            cur = torch.cuda.current_stream()
            with torch.cuda.stream(s_in):
                if state['comp_done'][b] is not None:
                    s_in.wait_event(state['comp_done'][b])                   # x_in[b] no longer being read
                else:
                    s_in.wait_stream(cur)
                t0 = torch.cuda.Event(enable_timing=True); t0.record(s_in)
                if world == 1:
                    x_in[b].copy_(x_host, non_blocking=True)                 # H2D of this step's input
                    t1 = torch.cuda.Event(enable_timing=True); t1.record(s_in)
                    t2 = t1
                else:
                    # every rank pulls 1/N of the input over ITS PCIe link, NVLink replicates it (V->E needs all rows)
                    if x_in_rep is not None:
                        x_in_rep[b].barrier()                                # every rank is done reading buffer b (step k-2)
                    x_in[b][sh.v_lo:sh.v_hi].copy_(x_host_mine, non_blocking=True)
                    t1 = torch.cuda.Event(enable_timing=True); t1.record(s_in)
                    if x_in_rep is not None:
                        _lib.push_rows(x_in[b][sh.v_lo:sh.v_hi], x_in_rep[b].peer_ptrs(sh.v_lo, unicast=True))
                        x_in_rep[b].barrier()                                # every rank's rows have landed everywhere
                    else:
                        sharding.allgather_rows(x_in[b], sh.v_ranges, rank)
                    t2 = torch.cuda.Event(enable_timing=True); t2.record(s_in)
                if _marks is not None:
                    ph_ev['h2d'].append((t0, t1)); ph_ev['gather'].append((t1, t2))
                ev = torch.cuda.Event(); ev.record(s_in); state['in_done'][b] = ev
            cur.wait_event(state['in_done'][b])
            if state['out_done'][b] is not None:
                cur.wait_event(state['out_done'][b])                         # result buffer b has been read back
            c0 = torch.cuda.Event(enable_timing=True); c0.record(cur)
            if world == 1:
                xe = allset_b200.segment_reduce(x_in[b], inc_v2e, None, 'sum')   # the call a user makes
                res = allset_b200.segment_reduce(xe, inc_e2v, None, 'sum')
                res.record_stream(s_out)
            else:
                sh.layer_pair_sum(x_in[b], x_e, xv_out[b], replicate_v=False)
                res = xv_out[b][sh.v_lo:sh.v_hi]                             # each rank reads back the rows it owns
            ev = torch.cuda.Event(enable_timing=True); ev.record(cur); state['comp_done'][b] = ev
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev)
                d0 = torch.cuda.Event(enable_timing=True); d0.record(s_out)
                out_host[b].copy_(res, non_blocking=True)                    # D2H of the result
                ev2 = torch.cuda.Event(enable_timing=True); ev2.record(s_out); state['out_done'][b] = ev2
            if _marks is not None:
                ph_ev['compute'].append((c0, ev)); ph_ev['d2h'].append((d0, ev2))
            state['k'] = k + 1

        def e2e_run(steps, warmup):
            for _ in range(warmup):
                e2e_step(None)
            torch.cuda.current_stream().wait_stream(s_out)
            barrier()
            start, end = ev(), ev()
            start.record()
            for _ in range(steps):
                e2e_step(True)
            torch.cuda.current_stream().wait_stream(s_out)                   # the last result must be on the host
            end.record()
            barrier()
            return max_over_ranks(start.elapsed_time(end))

        e2e_steps = max(4, min(a.steps, 20))
        e2e_ms = e2e_run(e2e_steps, 3)
        e2e_phases = {k: max_over_ranks(sum(a_.elapsed_time(b_) for a_, b_ in v) / max(len(v), 1)) for k, v in ph_ev.items()}
        e2e = {'value': Me / (e2e_ms / e2e_steps * 1e-3), 'unit': UNIT, 'phases_ms': e2e_phases, 'numa': numa,
               'h2d_bytes_per_step': int(Nv * d * es), 'd2h_bytes_per_step': int(Nv * d * es),
               'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps,
               'pipeline': 'H2D(k+1) || compute(k) || D2H(k-1), double-buffered, 3 streams' + ('' if world == 1 else
                            '; each rank copies 1/N of X_v from the host and replicates it over NVLink (%s)'
                            % ('P2P push into symmetric memory' if x_in_rep is not None else 'NCCL all-gather')),
               'api': 'allset_b200.segment_reduce(x, Incidence, None, "sum") x2' if world == 1
                      else 'allset_b200.sharding.ShardedIncidence.layer_pair_sum'}
        del x_host, x_host_mine, out_host, x_in, xv_out, x_in_rep
    clocks = sampler.stop()

    # ---- AllSetTransformer (PMA, heads=H) on the same graph: reported beside the headline -----------------------
    pma = None
    if not a.no_pma:
        g = torch.Generator(device=dev)
        g.manual_seed(a.seed + 1)
        score_v = torch.randn(Nv, H, device=dev, generator=g)
        score_e = torch.randn(Me, H, device=dev, generator=g)
        seed = torch.randn(H * (d // H), device=dev, generator=g)

        def pma_step(marks):
            if marks: marks[0].record()
            fe = sh.v2e_pma(x_v, score_v, seed, H, x_e)
            if marks: marks[1].record()
            sh.gather_e(x_e, fe)
            if marks: marks[2].record()
            fv = sh.e2v_pma(x_e, score_e, seed, H, x_v2 if replicate[0] else plain(x_v2))
            if marks: marks[3].record()
            if replicate[0]:
                sh.gather_v(x_v2, fv)
            if marks: marks[4].record()

        p_ms, pph = timed_steps(pma_step, 4, a.steps, a.warmup)
        nnz_e = int(sh.e_csr.nnz)
        nnz_v = int(sh.v_csr.nnz)
        b_ve = synthetic.algorithmic_bytes(nnz_e, sh.e_hi - sh.e_lo, d, es, heads=H)
        b_ev = synthetic.algorithmic_bytes(nnz_v, sh.v_hi - sh.v_lo, d, es, heads=H)
        pma = {'value': Me / (p_ms / a.steps * 1e-3), 'unit': UNIT, 'heads': H, 'ms_per_step': p_ms / a.steps,
               'v2e_ms': pph[0], 'e2v_ms': pph[2], 'exchange_x_e_ms': pph[1], 'exchange_x_v_ms': pph[3],
               'v2e_gbs': b_ve / (pph[0] * 1e-3) / 1e9, 'e2v_gbs': b_ev / (pph[2] * 1e-3) / 1e9}
        if world == 1 and H % 4 == 0:
            # V->E again with ONE packed [values | scores] record per vertex (what PMA.forward builds in bf16 mode when
            # the scores do not fit L2): same kernel, contiguous records
            from allset_b200 import _lib
            try:
                _, pv, ps = _lib.packed_pma_records(Nv, d, H, dtype, dev)
                pv.copy_(plain(x_v))
                ps.copy_(score_v)
                e = sh.e_csr

                def packed_step(marks):
                    if marks: marks[0].record()
                    _lib.pma_fwd_strided(pv, ps, seed, H, d // H, 0.2, e.rowptr, e.col, e.n_tgt)
                    if marks: marks[1].record()

                pk_ms, _ = timed_steps(packed_step, 1, max(5, a.steps // 5), a.warmup)
                pk_ms /= max(5, a.steps // 5)
                pma['v2e_packed_ms'] = pk_ms
                pma['v2e_packed_gbs'] = b_ve / (pk_ms * 1e-3) / 1e9
                del pv, ps
            except _lib.Unsupported as exc:
                pma['v2e_packed_ms'] = None
                pma['v2e_packed_note'] = str(exc)
        del score_v, score_e

    # ---- the dense glue of the same layer on the tensor cores (row (f)-1): f_enc / f_dec as ONE tcgen05 kernel ------
    mlp = None
    if not a.no_mlp and world == 1 and d in (64, 128):
        from allset_b200 import _lib
        gw = torch.Generator(device=dev)
        gw.manual_seed(a.seed + 2)
        w1 = torch.randn(d, d, device=dev, generator=gw) / d ** 0.5
        w2 = torch.randn(d, d, device=dev, generator=gw) / d ** 0.5
        bz = torch.zeros(d, device=dev)
        ln = (torch.ones(d, device=dev), torch.zeros(d, device=dev), 1e-5)
        x_rows = plain(x_v)
        o_dt = x_rows.dtype

        def mlp_step(marks):
            if marks: marks[0].record()
            _lib.mlp2_fwd(x_rows, w1, bz, w2, bz, ln, ln, True, o_dt)
            if marks: marks[1].record()

        m_ms, _ = timed_steps(mlp_step, 1, max(5, a.steps // 5), a.warmup)
        m_ms /= max(5, a.steps // 5)
        nbytes = 2 * Nv * d * es
        mlp = {'kernel': 'mlp2_ws_kernel (tcgen05, bf16 operands): relu(LN -> Linear -> ReLU -> LN -> Linear) over X_v',
               'rows': Nv, 'ms': m_ms, 'bytes': nbytes, 'gbs': nbytes / (m_ms * 1e-3) / 1e9,
               'tflops': 4.0 * Nv * d * d / (m_ms * 1e-3) / 1e12, 'rows_per_s': Nv / (m_ms * 1e-3)}
        # the training path's Linear kernels on the same rows (tcgen05): forward / input gradient / weight gradient, in the
        # bench dtype (bf16 operands) and for fp32 rows (three-term split precision, the reference's accuracy class)
        def lin_ms(fn):
            def step(marks):
                if marks: marks[0].record()
                fn()
                if marks: marks[1].record()
            n_it = max(5, a.steps // 5)
            t_ms, _ = timed_steps(step, 1, n_it, a.warmup)
            return t_ms / n_it

        lin_rows = min(Nv, 4_000_000)                 # fp32 copies of the rows: keep the extra footprint at 4 GB
        xr = x_rows[:lin_rows]
        dyr = torch.randn(lin_rows, d, device=dev, generator=gw).to(o_dt)
        linear = {'rows': lin_rows, 'kernels': 'allset_linear_fwd (mlp2_ws_kernel single-Linear modes) / allset_linear_wgrad (wgrad5::wgrad_kernel)'}
        for tag, xx, dd in (('bf16' if o_dt == torch.bfloat16 else 'f32_split', xr, dyr),) + \
                ((('f32_split', xr.float(), dyr.float()),) if o_dt == torch.bfloat16 else ()):
            io_b = 2 * lin_rows * d * xx.element_size()
            f_ms = lin_ms(lambda: _lib.linear_fwd(xx, w1))
            g_ms = lin_ms(lambda: _lib.linear_fwd(dd, w1, transposed=True))
            w_ms = lin_ms(lambda: _lib.linear_wgrad(dd, xx))
            linear[tag] = {'fwd_ms': f_ms, 'dgrad_ms': g_ms, 'wgrad_ms': w_ms, 'bytes_per_launch': io_b,
                           'fwd_gbs': io_b / (f_ms * 1e-3) / 1e9, 'dgrad_gbs': io_b / (g_ms * 1e-3) / 1e9,
                           'wgrad_gbs': io_b / (w_ms * 1e-3) / 1e9}
        mlp['linear'] = linear
        del w1, w2, xr, dyr

    # ---- roofline of the dominant kernel (the segmented-reduce gather kernel; both directions launch it) --------
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    nnz_e, nnz_v = int(sh.e_csr.nnz), int(sh.v_csr.nnz)
    b_ve = synthetic.algorithmic_bytes(nnz_e, sh.e_hi - sh.e_lo, d, es)
    b_ev = synthetic.algorithmic_bytes(nnz_v, sh.v_hi - sh.v_lo, d, es)
    t_ve, t_ge, t_ev, t_gv = ph
    achieved = (b_ve + b_ev) / ((t_ve + t_ev) * 1e-3) / 1e9
    roofline = {
        'bound': 'hbm', 'kernel': 'segreduce_stream_kernel<%s> (cp.async-staged segmented gather-reduce; both directions launch it)' % a.dtype,
        'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
        'peak_source': peak_src,
        'bytes_per_launch': (b_ve + b_ev) / 2, 'avg_launch_ms': (t_ve + t_ev) / 2,
        'v2e': {'bytes': b_ve, 'ms': t_ve, 'gbs': b_ve / (t_ve * 1e-3) / 1e9, 'frac': b_ve / (t_ve * 1e-3) / 1e9 / peak},
        'e2v': {'bytes': b_ev, 'ms': t_ev, 'gbs': b_ev / (t_ev * 1e-3) / 1e9, 'frac': b_ev / (t_ev * 1e-3) / 1e9 / peak},
    }
    traffic_path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if world == 1 and os.path.isfile(traffic_path):
        try:
            tr = json.load(open(traffic_path))
            if tr.get('workload') == workload_name(a):
                roofline['traffic'] = tr.get('dram_bytes_per_launch')
                roofline['traffic_source'] = 'static: ' + str(tr.get('source')) + ' (not re-measured in this run)'
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N == 1 only) -----------------------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        v, cores, desc, ms, done = cpu_time_pairs(a, steps=8, warmup=1, budget_s=20.0)
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc,
                        'ms_per_step_on_sample': ms, 'steps': done}
        try:
            cpu_baseline['half_layer_pair'] = cpu_time_half_layers(a)
        except Exception as exc:  # noqa  (reported, never fatal for the GPU line)
            cpu_baseline['half_layer_pair'] = {'error': repr(exc)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': a.dtype, 'data': 'synthetic',
            'config': {'workload': workload_name(a), 'nodes': Nv, 'hyperedges': Me, 'nnz': nnz, 'd': d,
                       'seed': a.seed, 'parallelism': 'single GPU' if world == 1 else
                       'hyperedge-sharded V->E / vertex-sharded E->V x%d; X_e exchanged between the directions every '
                       'step; updated X_v %s' % (world, 'replicated every step (another layer can follow)' if a.replicate_xv
                                                else 'left vertex-sharded (last layer: the next op is row-parallel); see other_mode'),
                       'exchange': exchange, 'push': os.environ.get('ALLSET_PUSH', 'direct (stores by the reducing warp)'),
                       'l2': 'inputs larger than L2 (X_v %.2f GB, col %.2f GB per step; no flush)'
                             % (Nv * d * es / 1e9, nnz * 4 / 1e9)},
            'clocks': clocks,
            'e2e': e2e,
            'gpu_launches': sh.launches_per_pair() * a.steps,
            'roofline': roofline,
            'cpu_baseline': cpu_baseline,
            'phases_ms': {'v2e': t_ve, 'exchange_x_e': t_ge, 'e2v': t_ev, 'exchange_x_v': t_gv},
            'other_mode': other, 'sharded_equals_unsharded': verified, 'sharded_check': verify_note,
            'v2e_only': {'value': Me / ((v2e_plain_ms if v2e_plain_ms else t_ve) * 1e-3), 'unit': UNIT,
                         'ms': v2e_plain_ms if v2e_plain_ms else t_ve,
                         'note': 'the V->E segmented reduce alone, without the fused X_e stores to the peers (max over ranks)'},
            'incidence_visits_per_s': 2 * nnz / (ms_per_step * 1e-3),
            'pma': pma,
            'mlp': mlp,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_b200(a)


if __name__ == '__main__':
    main()
