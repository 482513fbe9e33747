"""Debug aid (not a test): kernel-time breakdown of one eager MGCN training step via torch.profiler.

    python tests/step_profile.py [wn18rr|fb15k237|wikidata5m]     # big shapes use the fused loss (no dense [B, N] label)
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kgc_gcn_b200 as k
orc = bench.oracle()
dev = torch.device('cuda', 0)
workload = sys.argv[1] if len(sys.argv) > 1 else 'wn18rr'
N, R, E, seed = bench.WORKLOADS[workload]
tri = orc.synthetic_triples(N, R, E, seed)
g = orc.build_graph(tri, N, R)
prm = bench.params_ns()
graph = k.GraphData(edge_index=torch.from_numpy(g['edge_index']), edge_attr=torch.from_numpy(g['edge_attr']))
graph.entity = torch.from_numpy(g['entity']); graph.edge_norm = torch.from_numpy(g['edge_norm']); graph.num_nodes = N
graph.to(dev)
ds = k.KBDataset(bench.synthetic_query_set(k, tri, R), N, prm, training=True)
loader = k.BatchIterator(ds, bench.BATCH, shuffle=True, device=dev)
torch.manual_seed(0)
with torch.device(dev):
    model = k.MGCN(N, R, E, prm)
model.train()
opt = k.ClipAdam(model.parameters(), lr=1e-3, max_norm=1.0)
batches = loader.batches()
big = N * bench.BATCH * 4 > (1 << 30)
def step():
    qid = next(batches)
    opt.zero_grad()
    if big:
        loss = model.loss_sparse(qid, ds, graph)
    else:
        trip, lab = ds.build_batch(qid, dev)
        loss = model.loss(model(trip[:, 0], trip[:, 1], graph), lab)
    loss.backward()
    opt.step()
    return loss.item()
n_rep = 2 if big else 5
for _ in range(3): step()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(n_rep): step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / n_rep, e.count / n_rep) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('%s: total kernel time per step: %.1f us' % (workload, tot))
for name, us, n in rows[:45]:
    print('%9.1f us %5.1f%% x%-4.1f %s' % (us, 100 * us / tot, n, name[:120]))
