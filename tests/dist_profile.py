"""Debug aid (not a test): kernel-time breakdown of the dst-partitioned layer step on N GPUs (torchrun), rank 0 view.

    torchrun --nproc-per-node N tests/dist_profile.py            # strong scaling of the Wikidata5M shape (bench.py's N > 1 mode)
    KGC_PROFILE_WORKLOAD=wn18rr KGC_BENCH_WEAK=1 torchrun ...    # one WN18RR-shape partition per GPU
"""
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
workload = os.environ.get('KGC_PROFILE_WORKLOAD', 'wikidata5m')
scale = world if os.environ.get('KGC_BENCH_WEAK') else 1
case = bench.LayerCase(k, orc, workload, dev, world, rank, scale=scale)
part = case.part
if rank == 0:
    print('workload', workload, 'N', case.N, 'E', case.E, 'edges of rank 0', part.owned_eids.numel(), 'of', 2 * case.E, 'hubs', part.n_hub,
          'halo rows', 0 if getattr(part, 'halo_rows64', None) is None else part.halo_rows64.numel(), 'own rows', part.owned_nodes.numel())
for _ in range(5): case.step()
torch.cuda.synchronize(); dist.barrier()
# wall time of the eager step (events), max over ranks
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): case.step()
b.record(); torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / 5], device=dev, dtype=torch.float64)
ts = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(ts, t)
if rank == 0: print('eager step ms per rank:', ' '.join('%.2f' % float(v) for v in ts))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): case.step()
    torch.cuda.synchronize()
for r in range(world):
    if rank == r and r in (0, world - 1):
        rows = [(e.key, e.device_time_total / 5, e.count / 5) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == 'CUDA']
        rows.sort(key=lambda r: -r[1])
        print('rank %d: total kernel time per step: %.1f us' % (r, sum(r[1] for r in rows)))
        for name, us, n in rows[:32]:
            print('%9.1f us x%-4.1f %s' % (us, n, name[:110]))
        sys.stdout.flush()
    dist.barrier()
dist.barrier(); dist.destroy_process_group()
