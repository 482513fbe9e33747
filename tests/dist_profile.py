"""Debug aid (not a test): kernel-time breakdown of the dst-partitioned layer step on N GPUs (torchrun), rank 0 view."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kgc_gcn_b200 as k
orc = bench.oracle()
rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
dev = torch.device('cuda', lr); torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
N1, R, E1, seed = bench.WORKLOADS['wn18rr']
N, E = N1 * world, E1 * world
tri = orc.synthetic_triples(N, R, E, seed); g = orc.build_graph(tri, N, R)
p = orc.conv_params(N, R, E, 100, 200, seed=0)
torch.manual_seed(0)
conv = k.MGCNConv(100, 200, 2 * R).to(dev); conv.train()
part = k.GraphPartition(g['edge_index'], g['edge_attr'][0], N, 2 * R + 1, world, rank, dev, balance=os.environ.get('KGC_BALANCE', 'edges'))
if rank == 0: print('edges per rank', part.owned_eids.numel(), 'of', 2 * E, 'hubs', part.n_hub)
own = part.owned_nodes.cpu()
x = p['x'][own].to(dev).requires_grad_(True)
ee = p['edge_embs'][part.owned_eids.cpu()].to(dev).requires_grad_(True)
rl = p['rels'].to(dev).requires_grad_(True)
gen = torch.Generator().manual_seed(1)
g_ent = torch.randn(N, 200, generator=gen)[own].to(dev); g_rel = torch.randn(2 * R, 200, generator=gen).to(dev)
leaves = [x, ee, rl] + list(conv.parameters())
def step():
    for t in leaves: t.grad = None
    ent, rel = conv.forward_partitioned(x, part, ee, rl)
    torch.autograd.backward([ent, rel], [g_ent, g_rel])
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
if rank == 0:
    rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
    rows.sort(key=lambda r: -r[1])
    print('total kernel time per step: %.1f us' % sum(r[1] for r in rows))
    for name, us, n in rows[:30]:
        print('%8.1f us x%-4.1f %s' % (us, n, name[:110]))
dist.barrier(); dist.destroy_process_group()
