"""Debug aid (not a test): kernel-time breakdown of one eager MGCN training step (WN18RR shape) via torch.profiler."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kgc_gcn_b200 as k
orc = bench.oracle()
dev = torch.device('cuda', 0)
N, R, E, seed = bench.WORKLOADS['wn18rr']
tri = orc.synthetic_triples(N, R, E, seed)
g = orc.build_graph(tri, N, R)
prm = bench.params_ns()
graph = k.GraphData(edge_index=torch.from_numpy(g['edge_index']), edge_attr=torch.from_numpy(g['edge_attr']))
graph.entity = torch.from_numpy(g['entity']); graph.edge_norm = torch.from_numpy(g['edge_norm']); graph.num_nodes = N
graph.to(dev)
ds = k.KBDataset(bench.synthetic_queries(orc, tri, R), N, prm, training=True)
loader = k.BatchIterator(ds, bench.BATCH, shuffle=True, device=dev)
torch.manual_seed(0)
model = k.MGCN(N, R, E, prm).to(dev); model.train()
opt = k.ClipAdam(model.parameters(), lr=1e-3, max_norm=1.0)
batches = loader.batches()
def step():
    qid = next(batches)
    trip, lab = ds.build_batch(qid, dev)
    opt.zero_grad()
    pred = model(trip[:, 0], trip[:, 1], graph)
    loss = model.loss(pred, lab)
    loss.backward()
    opt.step()
    return loss.item()
for _ in range(5): step()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('total kernel time per step: %.1f us' % tot)
for name, us, n in rows[:45]:
    print('%8.1f us %5.1f%% x%-4.1f %s' % (us, 100 * us / tot, n, name[:120]))
