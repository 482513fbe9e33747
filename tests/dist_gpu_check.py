"""Multi-GPU check, launched by tests/test_gpu_dist.py as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_gpu_check.py
One process per GPU (NCCL).  (1) The dst-partitioned layer (all-gather of x, reduce-scatter of d_x, all-reduce of the
BatchNorm sums and of the replicated gradients) against the single-GPU layer on the whole graph; (2) the
entity-sharded filtered rank against the unsharded call - integer counts must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import mgcn_oracle as orc          # noqa: E402  (synthetic inputs only)
import kgc_gcn_b200 as k           # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', device_id=dev)
    N, R, E, D, Dout = 3000 * world + 1, 6, 9000 * world, 100, 200      # node count NOT a multiple of the rank count
    tri = orc.synthetic_triples(N, R, E, 77)
    g = orc.build_graph(tri, N, R)
    p = orc.conv_params(N, R, E, D, Dout, seed=3)
    gen = torch.Generator().manual_seed(4)
    g_ent, g_rel = torch.randn(N, Dout, generator=gen), torch.randn(2 * R, Dout, generator=gen)
    m_in = (torch.rand(N, Dout, generator=gen) > 0.1).to(torch.uint8)
    m_out = (torch.rand(N, Dout, generator=gen) > 0.1).to(torch.uint8)

    def make_conv():
        conv = k.MGCNConv(D, Dout, 2 * R, dropout=0.1).to(dev)
        with torch.no_grad():
            for name in ('loop_weight', 'in_weight', 'out_weight', 'rels_weight', 'loop_rel', 'loop_edge'):
                getattr(conv, name).copy_(p['w'][name])
        return conv.train()

    # ---- single-GPU answer on the whole graph (every rank computes it; deterministic kernels)
    ref = make_conv()
    ref.set_dropout_masks(m_in, m_out)
    x = p['x'].to(dev).requires_grad_(True)
    ee = p['edge_embs'].to(dev).requires_grad_(True)
    rl = p['rels'].to(dev).requires_grad_(True)
    ei, et = torch.from_numpy(g['edge_index']).to(dev), torch.from_numpy(g['edge_attr'][0]).to(dev)
    ent, rel = ref(x, ei, et, None, ee, rl)
    torch.autograd.backward([ent, rel], [g_ent.to(dev), g_rel.to(dev)])

    # ---- partitioned layer: the edge-balanced partition with split hub rows, then the plain range partition
    def check(a, b, name, tol=2e-5):
        scale = float(b.abs().max()) + 1e-30
        err = float((a - b).abs().max()) / scale
        assert err < tol, '{} rank {}: {:.3e}'.format(name, rank, err)
        return err
    errs = {}
    for balance, frac, p2p in (('hybrid', 0.5, 'auto'), ('edges', 0.1, 'auto'), ('edges', 0.5, 'auto'), ('edges', 0.1, False), ('range', 0.5, False)):
        part = k.GraphPartition(g['edge_index'], g['edge_attr'][0], N, 2 * R + 1, world, rank, dev, balance=balance,
                                hub_fraction=frac, p2p=p2p)
        if frac == 0.1 and balance == 'edges':
            assert part.n_hub >= 1                                 # the Zipf generator's hub must have been split
        own = part.owned_nodes
        conv = make_conv()
        conv.set_dropout_masks(m_in.to(dev)[own], m_out.to(dev)[own])
        xl = p['x'].to(dev)[own].clone().requires_grad_(True)
        eel = p['edge_embs'].to(dev)[part.owned_eids].clone().requires_grad_(True)
        rll = p['rels'].to(dev).requires_grad_(True)
        ent_l, rel_l = conv.forward_partitioned(xl, part, eel, rll)
        torch.autograd.backward([ent_l, rel_l], [g_ent.to(dev)[own], g_rel.to(dev)])
        e = {
            'all_ent': check(ent_l, ent[own], 'all_ent'), 'all_rel': check(rel_l, rel, 'all_rel'),
            'd_x': check(xl.grad, x.grad[own], 'd_x'), 'd_ee': check(eel.grad, ee.grad[part.owned_eids], 'd_ee'),
            'd_rel': check(rll.grad, rl.grad, 'd_rel'),
            'running_mean': check(conv.ent_bn.running_mean, ref.ent_bn.running_mean, 'running_mean'),
            'running_var': check(conv.ent_bn.running_var, ref.ent_bn.running_var, 'running_var'),
        }
        for name in ('loop_weight', 'in_weight', 'out_weight', 'rels_weight', 'loop_rel', 'loop_edge'):
            e['d_' + name] = check(getattr(conv, name).grad, getattr(ref, name).grad, 'd_' + name)
        e['d_gamma'] = check(conv.ent_bn.weight.grad, ref.ent_bn.weight.grad, 'd_gamma')
        e['d_beta'] = check(conv.ent_bn.bias.grad, ref.ent_bn.bias.grad, 'd_beta', tol=1e-3)   # ~0 in train mode: noise
        # evaluation mode, two different inputs back to back (no BatchNorm all-reduce between the two halo gathers)
        conv.eval(); ref.eval()
        with torch.no_grad():
            for shift in (0.0, 0.25):
                want = ref(x.detach() + shift, ei, et, None, ee.detach(), rl.detach())[0]
                got = conv.forward_partitioned(xl.detach() + shift, part, eel.detach(), rll.detach())[0]
                e['eval_%g' % shift] = check(got, want[own], 'eval all_ent')
        conv.train(); ref.train()
        n_own = torch.tensor([float(part.owned_eids.numel())], device=dev)
        dist.all_reduce(n_own, op=dist.ReduceOp.MAX)
        e['max_edge_share'] = float(n_own) / (2 * E)
        ctx = part.p2p(D)
        if ctx is not None:
            ctx.check()                                            # no barrier timed out
        e['p2p'] = 1.0 if ctx is not None else 0.0
        errs['{}/{}/{}'.format(balance, frac, p2p)] = e
        if balance in ('edges', 'hybrid'):
            assert e['max_edge_share'] <= 1.15 / world             # balanced: no rank owns much more than its share

    # ---- entity-sharded filtered rank
    B, NE, d = 300, 4096 * world, 200
    xq = torch.randint(-3, 4, (B, d), generator=gen).float().to(dev)
    tab = torch.randint(-3, 4, (NE, d), generator=gen).float().to(dev)
    bias = torch.randint(-2, 3, (NE,), generator=gen).float().to(dev)
    obj = torch.randint(0, NE, (B,), generator=gen).to(dev)
    fptr = torch.arange(0, 3 * B + 1, 3, dtype=torch.int64, device=dev)
    fidx = torch.randint(0, NE, (B, 3), generator=gen).sort(1).values.reshape(-1).to(torch.int32).to(dev)
    whole = k.filtered_rank(xq, tab, bias, obj, fptr, fidx, count_eq=True)
    per = NE // world
    shard = k.EntityTable(tab[rank * per:(rank + 1) * per].contiguous(), bias[rank * per:(rank + 1) * per].contiguous())
    sh = k.filtered_rank(xq, None, None, obj, fptr, fidx, count_eq=True, table=shard, n_offset=rank * per,
                         group=dist.group.WORLD)
    assert torch.equal(sh['ranks'], whole['ranks']) and torch.equal(sh['count_eq'], whole['count_eq'])
    assert torch.equal(sh['thr'], whole['thr'])
    dist.barrier()
    if rank == 0:
        print('DIST_OK', {b: {kk: float('{:.2e}'.format(v)) for kk, v in e.items()} for b, e in errs.items()})
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
