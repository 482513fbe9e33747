"""bench.py - headline benchmark of the M-GCN hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload wn18rr|fb15k237]

Metric (BASELINE.json): edges/sec of the relation-aware graph convolution, forward + backward, where an
"edge" is one directed edge of edge_index (2E per pass).  One JSON line on stdout:
  value        MGCNConv.forward + backward with every input resident in HBM (CUDA-event time, L2 flushed
               between steps, dropout p=0.1 drawn inside the timed region, CUDA-graph launch)
  e2e          the same metric through the reference-facing API for a whole training step
               (loader batch from host ids -> model(src, rel, graph) -> loss -> backward -> clip -> Adam ->
               loss.item()), host<->device copies inside the timed region
  roofline     the dominant kernel, timed alone with CUDA events: algorithmic bytes / time vs measured HBM peak
  cpu_baseline the oracle port of the reference (CPU torch, all host threads) on the same graph
  aux          filtered-rank queries/sec of the fused tcgen05 scorer (second half of BASELINE.json's metric)
`--impl reference` times the oracle port's full training step on the host cores (the reference is pure
Python and cannot travel to the GPU box; oracle/ is its pinned restatement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N entities, R relations, E train triples, synthetic seed)        SURVEY.md 8(d)
    'wn18rr': (40943, 11, 86835, 0),
    'fb15k237': (14541, 237, 272115, 1),
    'wikidata5m': (4594485, 822, 20614279, 2),
}
D_IN, D_OUT, BATCH = 100, 200, 128
# dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture of this bench command
# (profiles/r01_ncu_conv_kernels_lean.md); below the algorithmic bytes because x / g / rel rows hit L2 and part of the
# d_ee rows is still in L2 when the kernel ends
NCU_TRAFFIC = {'wn18rr': {'agg_fwd': 93.70e6, 'agg_bwd_src': 165.98e6, 'agg_bwd_rel': 111.75e6}}


def params_ns():
    from types import SimpleNamespace
    return SimpleNamespace(gcn_in_dim=D_IN, gcn_out_dim=D_OUT, gcn_drop=0.3, hidden_drop=0.3, feat_drop=0.3, k_w=10,
                           k_h=20, num_filter=200, kernel_size=7, bias=False, lbl_smooth=0.1, batch_size=BATCH)


def oracle():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import mgcn_oracle
    return mgcn_oracle


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def algorithmic_bytes(N, R, E, D=D_IN):
    """compulsory traffic, each tensor once, fp32, int32 indices."""
    T = 2 * R + 1
    fwd = 2 * E * (4 * D + 12) + 2 * N * (4 * D + 4) + N * 4 * D + T * 4 * D
    bwd = 2 * E * (8 * D + 12) + 2 * N * 4 * D + N * 4 * D + N * 4 * D + 2 * T * 4 * D
    return fwd, bwd


class ClockSampler(object):
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable'], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


def synthetic_queries(orc, tri, R):
    """(s, r) -> objects grouping of the train triples in both directions (data_loader.py:86-102)."""
    from collections import OrderedDict
    known = OrderedDict()
    for s, r, o in tri.tolist():
        known.setdefault((s, r), set()).add(o)
        known.setdefault((o, r + R), set()).add(s)
    return [{'triple': (s, r, -1), 'label': sorted(v)} for (s, r), v in known.items()]


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_conv_baseline(orc, N, R, E, seed, budget_s=12.0):
    """Oracle port of MGCNConv fwd+bwd (reference operation order) on the host cores; bounded sample."""
    torch.set_num_threads(os.cpu_count() or 1)
    tri = orc.synthetic_triples(N, R, E, seed)
    g = orc.build_graph(tri, N, R)
    p = orc.conv_params(N, R, E, D_IN, D_OUT, seed=0)
    ei, et = torch.from_numpy(g['edge_index']), torch.from_numpy(g['edge_attr'][0])
    gen = torch.Generator().manual_seed(1)
    g_ent, g_rel = torch.randn(N, D_OUT, generator=gen), torch.randn(2 * R, D_OUT, generator=gen)
    times = []
    t_end = time.time() + budget_s
    it = 0
    while it < 2 or (time.time() < t_end and it < 40):
        m_in = torch.empty(N, D_OUT).bernoulli_(0.9)
        m_out = torch.empty(N, D_OUT).bernoulli_(0.9)
        t0 = time.perf_counter()
        orc.conv_fwd_bwd(p['x'], ei, et, p['edge_embs'], p['rels'], p['w'], g_ent, g_rel, mask_in=m_in, mask_out=m_out)
        times.append(time.perf_counter() - t0)
        it += 1
    t = float(np.median(times[1:]))
    return {'value': 2 * E / t, 'unit': 'edges/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': 'oracle port of MGCNConv fwd+bwd (reference op order, fp32, dropout masks pre-drawn), full '
                      'graph, median of {} steps, {:.3f} s/step'.format(len(times) - 1, t)}


def run_reference(args, rank, world):
    """Reference arm: the oracle port's FULL training step on the host cores (same scope as our e2e)."""
    if rank != 0:
        return
    orc = oracle()
    N, R, E, seed = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    tri = orc.synthetic_triples(N, R, E, seed)
    g = orc.build_graph(tri, N, R)
    prm = params_ns()
    qs = synthetic_queries(orc, tri, R)
    model = orc.OracleMGCN(N, R, E, prm)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    ei, et = torch.from_numpy(g['edge_index']), torch.from_numpy(g['edge_attr'][0])
    rng = np.random.default_rng(0)
    times = []
    for i in range(args.warmup + args.steps):
        qid = rng.integers(0, len(qs), BATCH)
        t0 = time.perf_counter()
        trip, lab = orc.make_batch(qs, qid, N, prm.lbl_smooth, True)          # KBDataset + collate (data_loader.py:25-51)
        trip, lab = torch.from_numpy(trip), torch.from_numpy(lab)
        opt.zero_grad()
        pred = model(trip[:, 0], trip[:, 1], ei, et)
        loss = torch.nn.functional.binary_cross_entropy(pred, lab)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        loss.item()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    total = float(np.sum(times))
    val = 2 * E * len(times) / total
    line = {
        'impl': 'reference', 'metric': 'edges/sec GCN fwd+bwd', 'value': val, 'unit': 'edges/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload + '_shape', 'N': N, 'R': R, 'E': E, 'd_in': D_IN, 'd_out': D_OUT,
                   'batch': BATCH, 'scope': 'full training step on the host CPU'},
        'cpu_baseline': {'value': val, 'unit': 'edges/s', 'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': 'oracle port (oracle/mgcn_oracle.py OracleMGCN) full train step: label build, GCN, '
                                   'ConvE, BCE, backward, clip, Adam; whole graph each step'},
        'e2e': {'value': val, 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def time_kernel(fn, flush, iters=20, warm=3):
    for _ in range(warm):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def run_ours(args, rank, world, local_rank):
    import kgc_gcn_b200 as k
    L = k._lib
    L.lib()
    orc = oracle()                       # synthetic-workload generators + the cpu_baseline leg only
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    N1, R, E1, seed = WORKLOADS[args.workload]
    # N GPUs: weak scaling of the dst-partitioned layer - the global graph has world x the nodes and
    # triples of the named shape, nodes are range-partitioned by destination, every step all-gathers x (halo), reduce-
    # scatters d_x and all-reduces the BatchNorm sums and the replicated-parameter gradients over NCCL.
    N, E = N1 * world, E1 * world
    tri = orc.synthetic_triples(N, R, E, seed)
    g = orc.build_graph(tri, N, R)
    p = orc.conv_params(N, R, E, D_IN, D_OUT, seed=0)
    torch.manual_seed(0)
    conv = k.MGCNConv(D_IN, D_OUT, 2 * R).to(dev)          # dropout p = 0.1, the reference default (model.py:49)
    with torch.no_grad():
        for name in ('loop_weight', 'in_weight', 'out_weight', 'rels_weight', 'loop_rel', 'loop_edge'):
            getattr(conv, name).copy_(p['w'][name])
    conv.train()
    gen = torch.Generator().manual_seed(1)
    g_ent_all, g_rel = torch.randn(N, D_OUT, generator=gen), torch.randn(2 * R, D_OUT, generator=gen).to(dev)
    rl = p['rels'].to(dev).requires_grad_(True)
    if world == 1:
        part = None
        ei, et = torch.from_numpy(g['edge_index']).to(dev), torch.from_numpy(g['edge_attr'][0]).to(dev)
        x = p['x'].to(dev).requires_grad_(True)
        ee = p['edge_embs'].to(dev).requires_grad_(True)
        g_ent = g_ent_all.to(dev)
    else:
        part = k.GraphPartition(g['edge_index'], g['edge_attr'][0], N, 2 * R + 1, world, rank, dev,
                                p2p=False if os.environ.get('KGC_P2P', '1') == '0' else 'auto')      # KGC_P2P=0: NCCL halo exchange (A/B)
        own = part.owned_nodes.cpu()                     # edge-balanced partition: this rank's node rows (local-row order)
        x = p['x'][own].to(dev).requires_grad_(True)
        ee = p['edge_embs'][part.owned_eids.cpu()].to(dev).requires_grad_(True)
        g_ent = g_ent_all[own].to(dev)
    del g_ent_all
    leaves = [x, ee, rl] + list(conv.parameters())

    def step():
        for t in leaves:
            t.grad = None
        if part is None:
            ent, rel = conv(x, ei, et, None, ee, rl)
        else:
            ent, rel = conv.forward_partitioned(x, part, ee, rl)
        torch.autograd.backward([ent, rel], [g_ent, g_rel])

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_src = torch.zeros(64 << 20, dtype=torch.float32, device=dev)

    def flush():
        # write a buffer larger than L2 (126 MB), then stream a second one through it so that the dirty lines are
        # written back BEFORE the timed region starts (otherwise their eviction is charged to the timed kernel)
        flush_buf.zero_()
        flush_src.sum()

    # ---- launch mode: CUDA graph of the whole fwd+bwd (falls back to eager launches if capture fails)
    launch_mode = 'eager'
    graph = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    L.LAUNCHES = 0
    step()
    launches_per_step = L.LAUNCHES
    run_step = step
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            for t in leaves:
                t.grad = None
            with torch.cuda.graph(graph):
                step()
            run_step = graph.replay
            launch_mode = 'cuda_graph'
        except Exception as exc:                      # pragma: no cover
            sys.stderr.write('graph capture failed, timing eager launches: {}\n'.format(exc))
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        run_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        for a, b in ev:
            flush()
            a.record()
            run_step()
            b.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
        # ---- e2e: the whole training step through the reference-facing API, host ids in, loss out
        if args.no_e2e:
            e2e = {'ms_total': float('nan'), 'steps': 0, 'h2d': 0, 'd2h': 0, 'scope': 'skipped'}
        elif world == 1:
            e2e = e2e_train_step(k, orc, tri, g, N, R, E, dev, args)
        else:
            try:
                e2e = e2e_partitioned_step(conv, part, x, ee, rl, leaves, N, R, dev, args, dist)
            except Exception as exc:                  # pragma: no cover - keep the headline line if the e2e leg fails
                sys.stderr.write('partitioned e2e leg failed: {!r}\n'.format(exc))
                e2e = {'ms_total': float('nan'), 'steps': 0, 'h2d': 0, 'd2h': 0, 'scope': 'failed: {!r}'.format(exc)}
    if dist is not None:
        t = torch.tensor([total_ms, e2e['ms_total']], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e['ms_total'] = float(t[0]), float(t[1])
    value = 2 * E * args.steps / (total_ms * 1e-3)          # E already counts every rank's triples
    e2e_value = 2 * E * e2e['steps'] / (e2e['ms_total'] * 1e-3) if e2e['steps'] else None

    # ---- roofline of the dominant kernel, timed alone through the C ABI (rank 0)
    roof, kernels = None, None
    if rank == 0 and world == 1:
        roof, kernels = kernel_rooflines(k, conv, x, ee, rl, ei, et, N, R, E, flush, args.workload)
    if part is not None and part.p2p(D_IN) is not None:
        part.p2p(D_IN).check()                # a peer-memory barrier that timed out would have produced garbage timings
    if dist is not None:
        # release the CUDA graph (it holds the captured NCCL kernels) before the communicator goes away
        run_step = None
        graph = None
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        shutdown(dist)
        return
    fwd_b, bwd_b = algorithmic_bytes(N1, R, E1)          # per GPU
    peak, peak_src = measured_peaks()
    line = {
        'metric': 'edges/sec GCN fwd+bwd', 'value': value, 'unit': 'edges/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload + '_shape' + ('' if world == 1 else ' x{} (one shape-sized partition per GPU)'.format(world)), 'N': N, 'R': R, 'E': E, 'directed_edges': 2 * E, 'd_in': D_IN,
                   'd_out': D_OUT, 'dropout': 'p=0.1 keep masks drawn inside the timed region (training mode)',
                   'l2': 'flushed between steps (256 MiB memset + 256 MiB read, outside the timed events)', 'launch': launch_mode,
                   'parallelism': 'single GPU' if world == 1 else 'edge-balanced dst partition over {} GPUs ({} split hub rows); halo exchange of x / d_x: {}; all-reduce of hub rows, BN sums, replicated grads: NCCL'.format(world, part.n_hub, 'pulls over NVLink peer memory (K10)' if part.p2p(D_IN) is not None else 'NCCL all-gather / reduce-scatter')},
        'e2e': {'value': e2e_value, 'unit': 'edges/s', 'h2d_bytes_per_step': e2e['h2d'], 'd2h_bytes_per_step': e2e['d2h'],
                'ms_per_step': e2e['ms_total'] / e2e['steps'] if e2e['steps'] else None, 'eager_ms_per_step': e2e.get('eager_ms_per_step'),
                'dense_label_ms_per_step': e2e.get('dense_label_ms_per_step'), 'fused_loss_ms_per_step': e2e.get('fused_loss_ms_per_step'),
                'scope': e2e['scope']},
        'gpu_launches': launches_per_step * args.steps,
        'clocks': clocks.summary(),
        'roofline': roof,
        'kernels': kernels,
        'step_hbm': {'algorithmic_bytes_fwd_bwd': fwd_b + bwd_b, 'achieved_gbs': (fwd_b + bwd_b) / (total_ms / args.steps * 1e-3) / 1e9,
                     'frac_of_peak': (fwd_b + bwd_b) / (total_ms / args.steps * 1e-3) / 1e9 / peak, 'peak_gbs': peak,
                     'peak_source': peak_src, 'note': 'whole layer incl. the dense GEMMs and the BN/tanh tail, SURVEY.md 8(d) bytes'},
    }
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_conv_baseline(orc, N1, R, E1, seed)
    if not args.no_aux and world == 1:
        try:
            line['aux'] = aux_filtered_rank(k, dev, args)
        except Exception as exc:                       # pragma: no cover
            line['aux'] = {'error': repr(exc)}
    emit(line)
    shutdown(dist)


def shutdown(dist):
    """Tear the process group down; a watchdog ends the process if NCCL teardown stalls (the JSON line is already out)."""
    if dist is None:
        return
    sys.stdout.flush()
    timer = threading.Timer(20.0, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        dist.destroy_process_group()
    finally:
        timer.cancel()


def e2e_train_step(k, orc, tri, g, N, R, E, dev, args):
    """Whole training step through the public API, host ids in, loss out.  Two launch modes are timed: the graph-captured
    step (kgc_gcn_b200.GraphedTrainStep: one CUDA-graph replay per step - the headline e2e) and the plain eager loop the
    reference's main.py runs (reported as eager_ms_per_step).  The graph-captured step is timed with the dense [B, N] label
    (K5 + BCELoss) and with fused_loss=True (sparse positives; same loss and gradients); the headline is the
    faster of the two and both are reported."""
    prm = params_ns()
    graph = k.GraphData(edge_index=torch.from_numpy(g['edge_index']), edge_attr=torch.from_numpy(g['edge_attr']))
    graph.entity = torch.from_numpy(g['entity'])
    graph.edge_norm = torch.from_numpy(g['edge_norm'])
    graph.num_nodes = N
    graph.to(dev)
    ds = k.KBDataset(synthetic_queries(orc, tri, R), N, prm, training=True)
    loader = k.BatchIterator(ds, BATCH, shuffle=True, device=dev)
    steps, warm = args.steps, max(3, args.warmup)
    out = {}
    for mode in ('graph', 'eager', 'graph_fused'):
        torch.manual_seed(0)
        model = k.MGCN(N, R, E, prm).to(dev)
        model.train()
        opt = k.ClipAdam(model.parameters(), lr=1e-3, max_norm=1.0)          # clip_grad_norm_ + Adam (K9), main.py:68-71
        batches = loader.batches()
        step = k.GraphedTrainStep(model, opt, graph, ds, BATCH, fused_loss=(mode == 'graph_fused')) if mode != 'eager' else None
        ms, n = 0.0, 0
        try:
            for i in range(warm + steps):
                qid = next(batches)
                while len(qid) != BATCH:
                    qid = next(batches)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if step is not None:
                    step(qid).item()                                          # H2D: query ids; D2H: the loss
                else:
                    trip, lab = ds.build_batch(qid, dev)
                    opt.zero_grad()
                    pred = model(trip[:, 0], trip[:, 1], graph)
                    loss = model.loss(pred, lab)
                    loss.backward()
                    opt.step()                                                # clips inside (max_norm = 1.0)
                    loss.item()
                b.record()
                torch.cuda.synchronize()
                if i >= warm:
                    ms += a.elapsed_time(b)
                    n += 1
            out[mode] = (ms, n)
        except Exception as exc:                                              # pragma: no cover
            sys.stderr.write('e2e mode {} failed: {!r}\n'.format(mode, exc))
            torch.cuda.synchronize()
        del model, opt, step
        torch.cuda.empty_cache()
    per = {m: v[0] / v[1] for m, v in out.items() if v[1]}
    graphed = [m for m in ('graph', 'graph_fused') if m in per]
    head = min(graphed, key=per.get) if graphed else 'eager'
    how = {'graph': 'GraphedTrainStep, one CUDA-graph replay per step): host query ids (pinned staging buffer) -> K5 dense label build -> MGCN forward -> BCE',
           'graph_fused': 'GraphedTrainStep(fused_loss=True), one CUDA-graph replay per step): host query ids (pinned staging buffer) -> MGCN encoder + ConvE '
                          '-> 1-N scores -> BCE against the sparse positives fused with the logit gradient (N1)',
           'eager': 'eager loop): host query ids -> K5 dense label build -> MGCN forward -> BCE'}[head]
    res = {'ms_total': out[head][0], 'steps': out[head][1], 'h2d': BATCH * 8, 'd2h': 4,
           'scope': 'full training step through the public API (' + how + ' -> backward -> ClipAdam (gradient-norm clip + Adam, '
                    'K9) -> loss.item()',
           'dense_label_ms_per_step': per.get('graph'), 'fused_loss_ms_per_step': per.get('graph_fused')}
    if 'eager' in per and head != 'eager':
        res['eager_ms_per_step'] = per['eager']
    return res


def e2e_partitioned_step(conv, part, x, ee, rl, leaves, N, R, dev, args, dist):
    """N > 1: the partitioned encoder step through forward_partitioned with the batch coming from pinned host memory
    every step (H2D) and the scalar loss read back (D2H).  The loss is a stand-in decoder: mean of the batch's entity and
    relation rows, summed over ranks."""
    steps, warm = args.steps, max(3, args.warmup)
    rng = np.random.default_rng(0)
    host = torch.empty((BATCH, 2), dtype=torch.int64).pin_memory()
    local_row = torch.full((N,), -1, dtype=torch.int64, device=dev)
    local_row[part.owned_nodes] = torch.arange(part.n_loc, device=dev)
    ms, n = 0.0, 0
    for i in range(warm + steps):
        host[:, 0] = torch.from_numpy(rng.integers(0, N, BATCH))
        host[:, 1] = torch.from_numpy(rng.integers(0, 2 * R, BATCH))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        batch = host.to(dev, non_blocking=True)
        for t in leaves:
            t.grad = None
        ent, rel = conv.forward_partitioned(x, part, ee, rl)
        rows = local_row[batch[:, 0]]                    # -1: the entity lives on another rank
        mine = rows >= 0
        loss = (ent[rows.clamp_min(0)].mean(1) * mine).sum() + rel[batch[:, 1]].mean() / part.world
        loss.backward()
        tot = loss.detach().clone()
        dist.all_reduce(tot)
        tot.item()
        b.record()
        torch.cuda.synchronize()
        if i >= warm:
            ms += a.elapsed_time(b)
            n += 1
    return {'ms_total': ms, 'steps': n, 'h2d': BATCH * 16, 'd2h': 4,
            'scope': 'partitioned encoder step: batch ids from pinned host memory -> forward_partitioned -> scalar loss -> backward -> loss.item()'}


def kernel_rooflines(k, conv, x, ee, rl, ei, et, N, R, E, flush, workload='wn18rr'):
    """Each level-0 aggregation kernel alone through the C ABI, CUDA events, L2 flushed before every launch."""
    L = k._lib
    p, st = L.ptr, L.stream
    plan = k.get_plan(ei, et, N, 2 * R + 1)
    D, T = D_IN, 2 * R + 1
    relp = torch.cat([rl.detach(), conv.loop_rel.detach()], 0).contiguous()
    xd, eed = x.detach(), ee.detach()
    agg = torch.empty((2, N, D), device=x.device)
    g3 = torch.randn((3, N, D), device=x.device)
    d_ee, d_x, d_rel = torch.empty_like(eed), torch.empty((N, D), device=x.device), torch.empty((T, D), device=x.device)

    def carry(sp):
        return torch.empty((max(sp.n_carry, 1), D), device=x.device)
    pf, ps, pr = carry(plan.fwd), carry(plan.bwd_src), carry(plan.bwd_rel)
    sf, ss, sr = plan.fwd, plan.bwd_src, plan.bwd_rel
    n_f = n_s = n_r = (2 * E + 31) // 32                   # chunks (one warp each)
    fns = {
        'agg_fwd': lambda: L.call('kgc_agg_fwd', p(xd), p(relp), T, p(eed), p(plan.rec_dst), p(sf.rowflags), p(sf.chunks),
                                  sf.n_rec, p(agg), p(pf), D, st()),
        'agg_bwd_src': lambda: L.call('kgc_agg_bwd_src', p(xd), p(relp), T, p(eed), p(g3), p(plan.rec_src), p(ss.rowflags),
                                      p(ss.chunks), ss.n_rec, N, E, p(g3[2]), p(d_ee), p(d_x), p(ps), D, st()),
        'agg_bwd_rel': lambda: L.call('kgc_agg_bwd_rel', p(xd), p(eed), p(g3), p(plan.rec_type), p(sr.rowflags),
                                      p(sr.chunks), sr.n_rec, N, E, p(d_rel), p(pr), D, st()),
    }
    row = 4 * D
    bytes_ = {   # per launch: edge-embedding stream + 20 bytes of record per edge + each dense operand once + outputs once
        'agg_fwd': 2 * E * (row + 20) + N * row + T * row + 2 * N * row + 8 * n_f,
        'agg_bwd_src': 2 * E * (2 * row + 20) + N * row + 3 * N * row + T * row + N * row + 8 * n_s,
        'agg_bwd_rel': 2 * E * (row + 20) + N * row + 2 * N * row + T * row + 8 * n_r,
    }
    peak, peak_src = measured_peaks()
    out = {}
    cp_dst = torch.empty_like(eed)
    ms_cp = time_kernel(lambda: cp_dst.copy_(eed), flush)
    cp_bytes = 2 * eed.numel() * 4
    out['copy_same_size_reference'] = {'ms': ms_cp, 'algorithmic_bytes': cp_bytes, 'achieved_gbs': cp_bytes / (ms_cp * 1e-3) / 1e9,
                                       'frac': cp_bytes / (ms_cp * 1e-3) / 1e9 / peak,
                                       'note': 'torch copy of the edge-embedding table, timed the same way: what a plain '
                                               'streaming kernel of this size reaches (launch latency + tail included)'}
    # what a RANDOM gather of rows of this size reaches: the aggregation reads every edge-embedding row (400 B) and a node
    # row per edge in sorted-by-endpoint order, i.e. as a permutation of the tables
    perm = torch.randperm(eed.shape[0], device=x.device)
    ms_g = time_kernel(lambda: torch.index_select(eed, 0, perm, out=cp_dst), flush)
    g_bytes = cp_bytes + perm.numel() * 8
    out['row_gather_same_size_reference'] = {'ms': ms_g, 'algorithmic_bytes': g_bytes, 'achieved_gbs': g_bytes / (ms_g * 1e-3) / 1e9,
                                             'frac': g_bytes / (ms_g * 1e-3) / 1e9 / peak,
                                             'note': 'torch.index_select of all edge-embedding rows in a random order: the LIBRARY gather of '
                                                     'the same rows (it takes the same time for sequential indices and for any row '
                                                     'width, tests/gather_probe.py history: bound by its per-row overhead, not by HBM)'}
    del cp_dst, perm
    for name, fn in fns.items():
        ms = time_kernel(fn, flush)
        gbs = bytes_[name] / (ms * 1e-3) / 1e9
        out[name] = {'ms': ms, 'algorithmic_bytes': bytes_[name], 'achieved_gbs': gbs, 'frac': gbs / peak}
    top = max((n for n in out if n.startswith('agg_')), key=lambda n: out[n]['ms'])
    roof = {'kernel': top, 'bound': 'hbm', 'achieved': out[top]['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
            'frac': out[top]['frac'], 'traffic': NCU_TRAFFIC.get(workload, {}).get(top), 'peak_source': peak_src,
            'how': 'kernel launched alone through the C ABI, CUDA events on the launch stream, L2 flushed before each launch'}
    return roof, out


def aux_filtered_rank(k, dev, args):
    """Second half of BASELINE.json's metric: filtered-rank queries/sec of the fused tcgen05 scorer
    (B = 65,536 queries, d = 200, N entities; X = |N(0,1)|, E ~ U(-1,1), bias ~ N(0,0.1),
    mean 4 filtered positives per query, seed 3)."""
    L = k._lib
    out = []
    sizes = [1000000] if not args.aux_full else [1000000, 2000000, 4594485]
    B, d = 65536, D_OUT
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
        pk = json.load(f)
    for N in sizes:
        g = torch.Generator(device=dev).manual_seed(3)
        xq = torch.randn(B, d, generator=g, device=dev).abs_()
        tab = torch.rand(N, d, generator=g, device=dev).mul_(2).sub_(1)
        bias = torch.randn(N, generator=g, device=dev).mul_(0.1)
        obj = torch.randint(0, N, (B,), generator=g, device=dev)
        fptr = torch.arange(0, 4 * B + 1, 4, device=dev, dtype=torch.int64)
        fidx = torch.randint(0, N, (B, 4), generator=g, device=dev).sort(1).values.reshape(-1).to(torch.int32)

        def whole():
            return k.filtered_rank(xq, tab, bias, obj, fptr, fidx)
        for _ in range(2):
            res = whole()
        torch.cuda.synchronize()
        reps = 3
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b_ in ev:
            a.record()
            res = whole()
            b_.record()
        torch.cuda.synchronize()
        ms_whole = float(np.mean([a.elapsed_time(b_) for a, b_ in ev]))
        # the sweep kernel alone (dominant): packed operands resident, thresholds given
        table = k.EntityTable(tab, bias)
        q16 = k.pack_queries(xq)
        thr = res['thr']
        gt = torch.zeros(B, dtype=torch.int32, device=dev)

        def sweep():
            L.call('kgc_score_rank', L.ptr(q16), L.ptr(table.data), B, N, table.kpad, L.ptr(thr), L.ptr(gt), None, L.stream())
        ms_sweep = time_kernel(sweep, lambda: None, iters=3, warm=1)
        flops = 2.0 * B * N * d
        tf = flops / (ms_sweep * 1e-3) / 1e12
        out.append({'B': B, 'N': N, 'd': d, 'queries_per_s': B / (ms_whole * 1e-3), 'ms_whole_call': ms_whole,
                    'ms_sweep_kernel': ms_sweep, 'sweep_tflops_unpadded': tf,
                    'frac_of_bf16_sustained': tf / pk['bf16_tflops_sustained'], 'frac_of_bf16_burst': tf / pk['bf16_tflops'],
                    'mean_rank': float(res['sums'][1] / res['sums'][0])})
        del tab, table, xq
        torch.cuda.empty_cache()
    return {'metric': 'filtered-rank queries/sec (fused tcgen05 scoring, bf16 operands, fp32 accumulate)', 'unit': 'queries/s',
            'value': out[0]['queries_per_s'], 'sweeps': out,
            'roofline': {'bound': 'tensor', 'achieved': out[-1]['sweep_tflops_unpadded'], 'peak': pk['bf16_tflops_sustained'],
                         'unit': 'TFLOP/s', 'frac': out[-1]['frac_of_bf16_sustained'],
                         'note': 'un-padded flops 2*B*N*200 / sweep-kernel CUDA-event time vs measured sustained cuBLAS bf16'}}


_REAL_STDOUT = None


def guard_stdout():
    """Everything libraries print on fd 1 (NCCL prints its version there) goes to stderr; the ONE JSON line is written
    to the saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + '\n').encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='wn18rr', choices=sorted(WORKLOADS))
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-aux', action='store_true', help='skip the filtered-rank scoring sweep')
    ap.add_argument('--aux-full', action='store_true', help='scoring sweep at N = 1M, 2M and 4,594,485')
    ap.add_argument('--no-e2e', action='store_true', help='profiling runs only: skip the end-to-end leg')
    args = ap.parse_args()
    guard_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
