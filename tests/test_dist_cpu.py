"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the dst-range partition, the halo all-gather /
reduce-scatter plumbing of the partitioned layer and the entity-sharded rank reduction.  The per-rank
arithmetic is stood in for by the oracle (numpy) - no kernel runs here; the CUDA kernels behind the same
plumbing are checked on 2 GPUs by tests/test_gpu_dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mgcn_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _agg_numpy(x_full, rel, ee_rows, src, dst, typ, norm, n_rows):
    out = np.zeros((n_rows, x_full.shape[1]))
    np.add.at(out, dst, norm[:, None] * x_full[src] * rel[typ] * ee_rows)
    return out


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from kgc_gcn_b200.partition import partition_edges
        from kgc_gcn_b200.conv import _Collectives
        N, R, E, D = 64, 3, 300, 8
        tri = orc.synthetic_triples(N, R, E, 5)
        g = orc.build_graph(tri, N, R)
        ei, et = g['edge_index'], g['edge_attr'][0]
        rng = np.random.default_rng(0)
        x = rng.standard_normal((N, D))
        ee = rng.standard_normal((2 * E, D))
        rel = rng.standard_normal((2 * R + 1, D))
        info = partition_edges(ei, et, N, world, rank)
        lo, hi = info['lo'], info['hi']
        coll = _Collectives(None, world, N)
        # ---- halo all-gather puts the row blocks back in rank order
        x_full = coll.all_gather_rows(torch.from_numpy(x[lo:hi].copy())).numpy()
        np.testing.assert_array_equal(x_full, x)
        # ---- ownership: every edge owned exactly once, in-half edges first, global degrees
        owned = info['owned_eids']
        assert (np.diff(owned) > 0).all() and (owned[:info['n_edges_in']] < E).all() and (owned[info['n_edges_in']:] >= E).all()
        counts = torch.zeros(2 * E, dtype=torch.int64)
        counts[torch.from_numpy(owned)] = 1
        dist.all_reduce(counts)
        assert int(counts.min()) == 1 and int(counts.max()) == 1
        np.testing.assert_array_equal(info['deg'][0], orc.half_degree(ei[:, :E], N))
        np.testing.assert_array_equal(info['deg'][1], orc.half_degree(ei[:, E:], N))
        # ---- forward: local aggregation over owned edges == the rank's rows of the global aggregation
        norm = np.concatenate([orc.compute_norm(ei[:, :E], N).numpy(), orc.compute_norm(ei[:, E:], N).numpy()]).astype(np.float64)
        for h, sl in ((0, slice(0, info['n_edges_in'])), (1, slice(info['n_edges_in'], None))):
            eids = owned[sl]
            loc = _agg_numpy(x_full, rel, ee[eids], info['src'][sl], info['dst'][sl], info['type'][sl], norm[eids], hi - lo)
            half = slice(0, E) if h == 0 else slice(E, 2 * E)
            glob = _agg_numpy(x, rel, ee[half], ei[0, half], ei[1, half], et[half], norm[half], N)
            np.testing.assert_allclose(loc, glob[lo:hi], rtol=1e-12, atol=1e-12)
        # ---- backward: source-row gradients of the owned edges, reduce-scattered to the row owners
        gsel = rng.standard_normal((N, D))                       # d agg (both halves share it here)
        part = np.zeros((N, D))
        np.add.at(part, info['src'], norm[owned][:, None] * gsel[info['dst'] + lo] * rel[info['type']] * ee[owned])
        mine = coll.reduce_scatter_rows(torch.from_numpy(part)).numpy()
        full = np.zeros((N, D))
        np.add.at(full, ei[0], norm[:, None] * gsel[ei[1]] * rel[et] * ee)
        np.testing.assert_allclose(mine, full[lo:hi], rtol=1e-10, atol=1e-12)
        # ---- edge-balanced partition with split hub rows (partition.py): renumbered ids, virtual rows, hub exchange
        from kgc_gcn_b200.partition import partition_edges_balanced
        b = partition_edges_balanced(ei, et, N, world, rank, hub_fraction=0.05)
        n_loc, n_hub, blk, n_halo = b['n_loc'], b['n_hub'], b['block'], b['n_halo']
        assert n_hub >= 1 and n_loc == N // world and (b['dst'] < blk).all() and (b['src'] < blk + n_halo).all()
        own_e, own_n = b['owned_eids'], b['owned_nodes']
        counts = torch.zeros(2 * E, dtype=torch.int64)
        counts[torch.from_numpy(own_e)] = 1
        dist.all_reduce(counts)
        assert int(counts.min()) == 1 and int(counts.max()) == 1          # every edge owned exactly once
        ncount = torch.zeros(N, dtype=torch.int64)
        ncount[torch.from_numpy(own_n)] = 1
        dist.all_reduce(ncount)
        assert int(ncount.min()) == 1 and int(ncount.max()) == 1          # every node owned exactly once
        most = torch.tensor([own_e.shape[0]])
        dist.all_reduce(most, op=dist.ReduceOp.MAX)
        assert int(most) <= 1.15 * 2 * E / world                          # edge-balanced
        halo = b['halo_rows'].astype(np.int64)
        assert (np.diff(halo) > 0).all() and ((halo < rank * blk) | (halo >= (rank + 1) * blk)).all()
        comp_ids = np.concatenate([np.arange(rank * blk, (rank + 1) * blk), halo])      # compact id -> gathered-layout id
        old_of = np.full(world * blk, -1)
        old_of[b['newid']] = np.arange(N)                                 # gathered-layout id -> node (virtual rows: -1)
        real = old_of[comp_ids] >= 0
        np.testing.assert_array_equal(b['deg'][:, real], info['deg'][:, old_of[comp_ids][real]])   # degrees follow the ids
        mine = np.nonzero(b['hub_owner'] == rank)[0]
        collb = _Collectives(None, world, N, n_hub, torch.from_numpy(mine), torch.from_numpy(b['hub_row'][mine]), None, rank,
                             torch.from_numpy(halo))
        xb = np.concatenate([x[own_n], np.zeros((n_hub, D))])             # block of n_loc real + n_hub virtual rows
        table = collb.gather_compact(torch.from_numpy(xb)).numpy()        # own block, then the halo rows
        np.testing.assert_array_equal(table[real], x[old_of[comp_ids][real]])
        n_in = b['n_edges_in']
        planes = np.zeros((2, blk, D))
        for h, sl in ((0, slice(0, n_in)), (1, slice(n_in, None))):
            eids = own_e[sl]
            planes[h] = _agg_numpy(table, rel, ee[eids], b['src'][sl], b['dst'][sl], b['type'][sl], norm[eids], blk)
        planes_t = torch.from_numpy(planes)
        collb.sum_hub_rows(planes_t, n_loc)
        for h in (0, 1):
            half = slice(0, E) if h == 0 else slice(E, 2 * E)
            glob = _agg_numpy(x, rel, ee[half], ei[0, half], ei[1, half], et[half], norm[half], N)
            np.testing.assert_allclose(planes_t[h, :n_loc].numpy(), glob[own_n], rtol=1e-10, atol=1e-11)
        # backward: upstream rows of the hubs reach every rank's virtual rows; d_x returns through the reduce-scatter
        g_loc = np.zeros((1, blk, D))
        g_loc[0, :n_loc] = gsel[own_n]
        g_t = torch.from_numpy(g_loc)
        collb.spread_hub_rows(g_t, 1, n_loc)
        np.testing.assert_array_equal(g_t[0, n_loc:].numpy(), gsel[b['hubs']])
        partb = np.zeros((blk + n_halo, D))
        np.add.at(partb, b['src'], norm[own_e][:, None] * g_t[0].numpy()[b['dst']] * rel[b['type']] * ee[own_e])
        mine_dx = collb.reduce_compact(torch.from_numpy(partb), blk).numpy()
        np.testing.assert_allclose(mine_dx[:n_loc], full[own_n], rtol=1e-10, atol=1e-12)
        # the owner-side index table of the peer-memory reduce (K10): the same sum, pulled row by row in rank order
        gathered = [torch.zeros((blk + b['n_halo_max'], D), dtype=torch.float64) for _ in range(world)]
        padded = torch.zeros((blk + b['n_halo_max'], D), dtype=torch.float64)
        padded[:blk + n_halo] = torch.from_numpy(partb)
        dist.all_gather(gathered, padded)
        pulled = np.zeros((n_loc, D))
        for r in range(world):
            at = b['peer_idx'][r]
            pulled[at >= 0] += gathered[r].numpy()[at[at >= 0]]
        np.testing.assert_allclose(pulled, full[own_n], rtol=1e-10, atol=1e-12)
        # ---- hybrid cut (partition_edges_hybrid): an edge lives with its lower-degree endpoint; remote rows are sources AND
        # destinations; partial aggregates go to their owners through peer_dst_idx, partial d_x through peer_idx
        from kgc_gcn_b200.partition import partition_edges_hybrid
        hb = partition_edges_hybrid(ei, et, N, world, rank)
        hl, hrem = hb['n_loc'], hb['n_halo']
        rows_h, rows_max = hl + hrem, hl + hb['n_halo_max']
        assert hb['block'] == hl and hb['n_hub'] == 0 and int(hb['n_remote_all'][rank]) == hrem
        he, hn = hb['owned_eids'], hb['owned_nodes']
        counts = torch.zeros(2 * E, dtype=torch.int64)
        counts[torch.from_numpy(he)] = 1
        dist.all_reduce(counts)
        assert int(counts.min()) == 1 and int(counts.max()) == 1          # every edge owned exactly once
        ncount = torch.zeros(N, dtype=torch.int64)
        ncount[torch.from_numpy(hn)] = 1
        dist.all_reduce(ncount)
        assert int(ncount.min()) == 1 and int(ncount.max()) == 1          # every node owned exactly once
        most = torch.tensor([he.shape[0]])
        dist.all_reduce(most, op=dist.ReduceOp.MAX)
        assert int(most) <= 1.15 * 2 * E / world                          # edge-balanced without split rows
        rem = hb['halo_rows'].astype(np.int64)
        assert (np.diff(rem) > 0).all() and ((rem < rank * hl) | (rem >= (rank + 1) * hl)).all()
        assert hrem < b['n_halo']                                         # fewer remote rows than the destination partition's halo
        old_h = np.full(world * hl, -1)
        old_h[hb['newid']] = np.arange(N)
        comp_h = np.concatenate([np.arange(rank * hl, (rank + 1) * hl), rem])
        valid = old_h[comp_h] >= 0
        assert valid[hl:].all() and (hb['src'] < rows_h).all() and (hb['dst'] < rows_h).all()
        # an edge's anchor (an endpoint this rank owns) is its lower-degree endpoint, ties to the destination
        tot = np.bincount(ei[0], minlength=N) + np.bincount(ei[1], minlength=N)
        anchor_is_src = tot[ei[0, he]] < tot[ei[1, he]]
        assert (np.where(anchor_is_src, hb['src'], hb['dst']) < hl).all()
        np.testing.assert_array_equal(hb['deg'][:, valid], info['deg'][:, old_h[comp_h][valid]])
        table_h = np.zeros((rows_h, D))
        table_h[valid] = x[old_h[comp_h][valid]]
        n_in_h = hb['n_edges_in']
        planes_h = np.zeros((2, rows_max, D))
        for h, sl in ((0, slice(0, n_in_h)), (1, slice(n_in_h, None))):
            eids = he[sl]
            planes_h[h, :rows_h] = _agg_numpy(table_h, rel, ee[eids], hb['src'][sl], hb['dst'][sl], hb['type'][sl], norm[eids], rows_h)
        everyone = [torch.zeros((2, rows_max, D), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(everyone, torch.from_numpy(planes_h))
        n_real_h = hb['n_real']
        assert (hb['peer_dst_idx'][rank] == np.arange(n_real_h)).all() and (hb['peer_idx'][rank] == np.arange(n_real_h)).all()
        for h in (0, 1):
            summed = np.zeros((n_real_h, D))
            for r in range(world):                                         # the owner adds the partial rows in rank order
                at = hb['peer_dst_idx'][r]
                summed[at >= 0] += everyone[r].numpy()[h][at[at >= 0]]
            half = slice(0, E) if h == 0 else slice(E, 2 * E)
            glob = _agg_numpy(x, rel, ee[half], ei[0, half], ei[1, half], et[half], norm[half], N)
            np.testing.assert_allclose(summed, glob[hn], rtol=1e-10, atol=1e-11)
        # backward: the upstream rows of the remote destinations come from their owners; partial d_x goes back through peer_idx
        g_tab = np.zeros((rows_h, D))
        g_tab[valid] = gsel[old_h[comp_h][valid]]
        part_h = np.zeros((rows_max, D))
        np.add.at(part_h, hb['src'], norm[he][:, None] * g_tab[hb['dst']] * rel[hb['type']] * ee[he])
        everyone = [torch.zeros((rows_max, D), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(everyone, torch.from_numpy(part_h))
        pulled = np.zeros((n_real_h, D))
        for r in range(world):
            at = hb['peer_idx'][r]
            pulled[at >= 0] += everyone[r].numpy()[at[at >= 0]]
        np.testing.assert_allclose(pulled, full[hn], rtol=1e-10, atol=1e-12)
        # ---- entity-sharded filtered rank: integer counts all-reduce to the unsharded answer, target logits sum exactly
        B, NE = 16, 64
        scores = rng.integers(-5, 6, (B, NE)).astype(np.float64)
        obj = rng.integers(0, NE, B)
        fptr = np.arange(0, 2 * B + 1, 2)
        fidx = rng.integers(0, NE, 2 * B)
        gt, eq = orc.rank_counts(scores, fptr, fidx, obj)
        per = NE // world
        s_lo, s_hi = rank * per, (rank + 1) * per
        thr_local = np.where((obj >= s_lo) & (obj < s_hi), scores[np.arange(B), obj], 0.0)
        thr = torch.from_numpy(thr_local)
        dist.all_reduce(thr)                                      # each entry non-zero on exactly one rank
        np.testing.assert_array_equal(thr.numpy(), scores[np.arange(B), obj])
        keep = np.ones((B, NE), dtype=bool)
        for q in range(B):
            keep[q, fidx[fptr[q]:fptr[q + 1]]] = False
            keep[q, obj[q]] = False
        loc_gt = torch.from_numpy(((scores[:, s_lo:s_hi] > thr.numpy()[:, None]) & keep[:, s_lo:s_hi]).sum(1))
        dist.all_reduce(loc_gt)
        np.testing.assert_array_equal(loc_gt.numpy(), gt)
        ret[rank] = 'ok'
    finally:
        dist.destroy_process_group()


def test_partition_and_collectives_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 'ok', 1: 'ok'}


@pytest.mark.parametrize('world', [2, 4, 8])
def test_partitions_with_uneven_node_counts(world):
    """num_nodes not a multiple of the number of ranks (the Wikidata5M shape: 4,594,485 nodes on 2 / 4 / 8 GPUs): blocks
    keep the stride ceil(N / world), some ranks hold one real row less, no edge refers to a row that does not exist, and
    the per-rank aggregation / owner-side gradient pull (emulated in numpy, one rank after the other) give the global
    answer."""
    from kgc_gcn_b200.partition import partition_edges, partition_edges_balanced
    N, R, E, D = 64 * world + 3, 3, 400 * world, 4
    tri = orc.synthetic_triples(N, R, E, 13)
    g = orc.build_graph(tri, N, R)
    ei, et = g['edge_index'], g['edge_attr'][0]
    rng = np.random.default_rng(1)
    x, ee, rel = rng.standard_normal((N, D)), rng.standard_normal((2 * E, D)), rng.standard_normal((2 * R + 1, D))
    gsel = rng.standard_normal((N, D))
    glob = _agg_numpy(x, rel, ee, ei[0], ei[1], et, np.ones(2 * E), N)
    full = np.zeros((N, D))
    np.add.at(full, ei[0], gsel[ei[1]] * rel[et] * ee)
    per = -(-N // world)
    # ---- range partition
    infos = [partition_edges(ei, et, N, world, r) for r in range(world)]
    assert sorted(np.concatenate([i['owned_eids'] for i in infos]).tolist()) == list(range(2 * E))
    assert [i['hi'] - i['lo'] for i in infos] == [min(per, max(0, N - r * per)) for r in range(world)]
    for r, i in enumerate(infos):
        assert i['per'] == per and i['deg'].shape == (2, world * per) and (i['deg'][:, N:] == 0).all()
        assert (i['dst'] >= 0).all() and (i['dst'] < i['hi'] - i['lo']).all()
        loc = _agg_numpy(x, rel, ee[i['owned_eids']], i['src'], i['dst'], i['type'], np.ones(i['owned_eids'].shape[0]), per)
        np.testing.assert_allclose(loc[:i['hi'] - i['lo']], glob[i['lo']:i['hi']], rtol=1e-12, atol=1e-12)
    # ---- edge-balanced partition with split hubs
    parts = [partition_edges_balanced(ei, et, N, world, r, hub_fraction=0.1) for r in range(world)]
    assert sorted(np.concatenate([p['owned_eids'] for p in parts]).tolist()) == list(range(2 * E))
    assert sorted(np.concatenate([p['owned_nodes'] for p in parts]).tolist()) == list(range(N))
    counts = [p['owned_nodes'].shape[0] for p in parts]
    assert max(counts) == per and min(counts) >= per - 1 and sum(counts) == N
    blk, n_loc, n_hub = parts[0]['block'], parts[0]['n_loc'], parts[0]['n_hub']
    assert n_loc == per and blk == per + n_hub
    tables, partials = [], []
    for r, p in enumerate(parts):
        assert p['n_real'] == counts[r] and p['peer_idx'].shape == (world, counts[r])
        comp_ids = np.concatenate([np.arange(r * blk, (r + 1) * blk), p['halo_rows'].astype(np.int64)])
        old_of = np.full(world * blk, -1)
        old_of[p['newid']] = np.arange(N)
        table = np.zeros((blk + p['n_halo'], D))
        real = old_of[comp_ids] >= 0
        table[real] = x[old_of[comp_ids][real]]
        assert real[p['src']].all()                                     # no edge reads a row that does not exist
        assert ((p['dst'] < counts[r]) | (p['dst'] >= n_loc)).all()      # ... or writes one
        own_e = p['owned_eids']
        planes = _agg_numpy(table, rel, ee[own_e], p['src'], p['dst'], p['type'], np.ones(own_e.shape[0]), blk)
        tables.append(planes)
        g_loc = np.zeros((blk, D))
        g_loc[:counts[r]] = gsel[p['owned_nodes']]
        g_loc[n_loc:] = gsel[p['hubs']]
        part = np.zeros((blk + p['n_halo_max'], D))
        np.add.at(part, p['src'], g_loc[p['dst']] * rel[p['type']] * ee[own_e])
        partials.append(part)
    hub_sum = sum(t[n_loc:] for t in tables)                            # virtual rows summed over ranks
    for r, p in enumerate(parts):
        out = tables[r][:counts[r]].copy()
        mine = np.nonzero(p['hub_owner'] == r)[0]
        out[p['hub_row'][mine]] = hub_sum[mine]
        np.testing.assert_allclose(out, glob[p['owned_nodes']], rtol=1e-10, atol=1e-11)
        pulled = np.zeros((counts[r], D))
        for q in range(world):
            at = p['peer_idx'][q]
            pulled[at >= 0] += partials[q][at[at >= 0]]
        np.testing.assert_allclose(pulled, full[p['owned_nodes']], rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize('world', [2, 3, 4, 8])
def test_hybrid_cut_all_ranks(world):
    """partition_edges_hybrid on the hub-heavy generator, every rank emulated in numpy (node counts not a multiple of the
    rank count): every edge and node owned once, edge-balanced without split rows, far fewer remote rows than the
    destination partition's halo, and the four exchanges (x rows in, partial aggregates to their owners through
    peer_dst_idx, upstream rows in, partial d_x to their owners through peer_idx) reproduce the global aggregation and the
    global source-row gradients; the sparse tables (sparse_peer_table) list exactly the rows other ranks contribute to."""
    from kgc_gcn_b200.partition import partition_edges_hybrid, partition_edges_balanced, sparse_peer_table
    N, R, E, D = 256 * world + 3, 3, 2000 * world, 4
    tri = orc.synthetic_triples(N, R, E, 17)
    g = orc.build_graph(tri, N, R)
    ei, et = g['edge_index'], g['edge_attr'][0]
    rng = np.random.default_rng(2)
    x, ee, rel = rng.standard_normal((N, D)), rng.standard_normal((2 * E, D)), rng.standard_normal((2 * R + 1, D))
    g_in, g_out = rng.standard_normal((N, D)), rng.standard_normal((N, D))
    glob = [_agg_numpy(x, rel, ee[h], ei[0, h], ei[1, h], et[h], np.ones(E), N) for h in (slice(0, E), slice(E, 2 * E))]
    full = np.zeros((N, D))
    np.add.at(full, ei[0, :E], g_in[ei[1, :E]] * rel[et[:E]] * ee[:E])
    np.add.at(full, ei[0, E:], g_out[ei[1, E:]] * rel[et[E:]] * ee[E:])
    parts = [partition_edges_hybrid(ei, et, N, world, r) for r in range(world)]
    assert sorted(np.concatenate([p['owned_eids'] for p in parts]).tolist()) == list(range(2 * E))
    assert sorted(np.concatenate([p['owned_nodes'] for p in parts]).tolist()) == list(range(N))
    assert max(p['owned_eids'].shape[0] for p in parts) <= 1.1 * 2 * E / world
    halo_dst = sum(partition_edges_balanced(ei, et, N, world, r, hub_fraction=0.1)['n_halo'] for r in range(world))
    assert sum(p['n_halo'] for p in parts) < 0.5 * halo_dst
    n_loc = parts[0]['n_loc']
    rows_max = n_loc + max(p['n_halo'] for p in parts)
    aggs, dxs = [], []
    for r, p in enumerate(parts):
        assert p['n_loc'] == n_loc == -(-N // world) and p['block'] == n_loc and p['n_hub'] == 0
        assert [int(v) for v in p['n_remote_all']] == [q['n_halo'] for q in parts] and p['n_halo_max'] == rows_max - n_loc
        rows = n_loc + p['n_halo']
        comp = np.concatenate([np.arange(r * n_loc, (r + 1) * n_loc), p['halo_rows'].astype(np.int64)])
        old = np.full(world * n_loc, -1)
        old[p['newid']] = np.arange(N)
        valid = old[comp] >= 0
        assert valid[p['src']].all() and valid[p['dst']].all() and (p['src'] < rows).all() and (p['dst'] < rows).all()
        xt, gi, go = np.zeros((rows, D)), np.zeros((rows, D)), np.zeros((rows, D))
        xt[valid], gi[valid], go[valid] = x[old[comp][valid]], g_in[old[comp][valid]], g_out[old[comp][valid]]   # pulled rows
        n_in, own = p['n_edges_in'], p['owned_eids']
        agg = np.zeros((2, rows_max, D))
        agg[0, :rows] = _agg_numpy(xt, rel, ee[own[:n_in]], p['src'][:n_in], p['dst'][:n_in], p['type'][:n_in], np.ones(n_in), rows)
        agg[1, :rows] = _agg_numpy(xt, rel, ee[own[n_in:]], p['src'][n_in:], p['dst'][n_in:], p['type'][n_in:],
                                   np.ones(own.shape[0] - n_in), rows)
        aggs.append(agg)
        dx = np.zeros((rows_max, D))
        np.add.at(dx, p['src'][:n_in], gi[p['dst'][:n_in]] * rel[p['type'][:n_in]] * ee[own[:n_in]])
        np.add.at(dx, p['src'][n_in:], go[p['dst'][n_in:]] * rel[p['type'][n_in:]] * ee[own[n_in:]])
        dxs.append(dx)
    for r, p in enumerate(parts):
        nr = p['n_real']
        for h in (0, 1):
            tot = np.zeros((nr, D))
            for q in range(world):
                at = p['peer_dst_idx'][q]
                tot[at >= 0] += aggs[q][h][at[at >= 0]]
            np.testing.assert_allclose(tot, glob[h][p['owned_nodes']], rtol=1e-10, atol=1e-11)
        tot = np.zeros((nr, D))
        for q in range(world):
            at = p['peer_idx'][q]
            tot[at >= 0] += dxs[q][at[at >= 0]]
        np.testing.assert_allclose(tot, full[p['owned_nodes']], rtol=1e-10, atol=1e-11)
        # the sparse form used on the device: own value in place, only the listed rows get the other ranks' partial rows
        for table, mine, want in ((p['peer_dst_idx'], aggs[r][0][:nr].copy(), glob[0][p['owned_nodes']]),):
            rows_l, idx_l = (a.numpy() for a in sparse_peer_table(table, r))
            assert (idx_l[r] == -1).all() and idx_l.shape == (world, rows_l.shape[0])
            for q in range(world):
                at = idx_l[q]
                mine[rows_l[at >= 0]] += aggs[q][0][at[at >= 0]]
            np.testing.assert_allclose(mine, want, rtol=1e-10, atol=1e-11)


@pytest.mark.parametrize('world', [2, 4, 8])
def test_balanced_partition_properties(world):
    """Host logic of the edge-balanced partition (partition.py) on the hub-heavy generator of SURVEY.md 8(d): equal node
    counts, every edge and node owned once, near-equal edge counts where the range partition is 1.5x - 4.5x off, hubs split
    only where no whole-row assignment could balance them, compact tables that hold exactly the rows the edges read."""
    from kgc_gcn_b200.partition import partition_edges, partition_edges_balanced
    N, R, E = 2048 * world, 5, 8192 * world
    tri = orc.synthetic_triples(N, R, E, 11)
    g = orc.build_graph(tri, N, R)
    ei, et = g['edge_index'], g['edge_attr'][0]
    parts = [partition_edges_balanced(ei, et, N, world, r) for r in range(world)]
    mean = 2 * E / world
    assert max(p['owned_eids'].shape[0] for p in parts) <= 1.1 * mean
    assert max(partition_edges(ei, et, N, world, r)['owned_eids'].shape[0] for r in range(world)) >= 1.4 * mean
    assert sorted(np.concatenate([p['owned_eids'] for p in parts]).tolist()) == list(range(2 * E))
    assert sorted(np.concatenate([p['owned_nodes'] for p in parts]).tolist()) == list(range(N))
    n_hub = parts[0]['n_hub']
    indeg = np.bincount(ei[1], minlength=N)
    assert n_hub == int((indeg > 0.5 * mean).sum()) and (n_hub > 0) == (indeg.max() > 0.5 * mean)
    for r, p in enumerate(parts):
        blk = p['block']
        assert p['n_loc'] == N // world and blk == N // world + n_hub and p['n_hub'] == n_hub
        assert p['src'].min() >= 0 and p['src'].max() < blk + p['n_halo'] and p['dst'].max() < blk
        used = np.unique(p['src'])
        assert (used[used >= blk] == blk + np.arange(p['n_halo'])).all()          # every halo row is read, none is missing
        assert p['n_halo'] <= p['n_halo_max'] == max(q['n_halo'] for q in parts)
        # the owner-side table of the peer-memory reduce: rank q reads my row v <=> peer_idx[q, v] is its compact row there
        for q, other in enumerate(parts):
            at = p['peer_idx'][q]
            ids_mine = r * blk + np.arange(p['n_loc'])                            # my rows in the gathered layout
            if q == r:
                read = np.isin(np.arange(p['n_loc']), other['src'][other['src'] < blk])
                assert ((at >= 0) == read).all() and (at[at >= 0] == np.arange(p['n_loc'])[at >= 0]).all()
            else:
                halo = other['halo_rows'].astype(np.int64)
                pos = np.searchsorted(halo, ids_mine)
                hit = (pos < halo.shape[0]) & (halo[np.minimum(pos, halo.shape[0] - 1)] == ids_mine)
                assert ((at >= 0) == hit).all() and (at[hit] == blk + pos[hit]).all()


def test_halo_pull_order_is_a_permutation_that_rotates_owners():
    """partition.halo_pull_order: every position once; consecutive 8-row trips go to different owners, starting with the
    owner behind the reader's rank (no owner is read by every rank at the same time)."""
    import numpy as np
    from kgc_gcn_b200.partition import halo_pull_order
    rng = np.random.default_rng(5)
    block, world = 1000, 8
    for rank in (0, 3, 7):
        remote = np.setdiff1d(np.arange(world * block), np.arange(rank * block, (rank + 1) * block))
        ids = np.sort(rng.choice(remote, 3001, replace=False))
        order = halo_pull_order(ids, block, rank, world)
        assert order.dtype == np.int32 and sorted(order.tolist()) == list(range(ids.shape[0]))
        owners = ids[order[order % 8 == 0]] // block
        first = owners[:world - 1].tolist()
        assert first == [(rank + 1 + j) % world for j in range(world - 1)]
        assert (owners[1:world - 1] != owners[:world - 2]).all()
        inner = np.nonzero(order % 8 != 0)[0]                            # a trip is 8 consecutive list positions
        assert (order[inner] == order[inner - 1] + 1).all()
    assert halo_pull_order(np.zeros((0,), dtype=np.int64), block, 0, world).shape == (0,)
