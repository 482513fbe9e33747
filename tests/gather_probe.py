"""Debug aid (not a test): how much of the aggregation kernel's time is the RANDOM order of its 400-byte row reads?
Hand-made dst-sorted records (rows of 4 records, no carries) with sequential / random edge ids and source rows."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgc_gcn_b200 as k
L = k._lib
p, st = L.ptr, L.stream

def timed(fn, reps=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(reps):
        flush.zero_(); flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]

n_rec, N, D, T = 173664, 40943, 100, 23
rng = np.random.default_rng(0)
x = torch.randn(N, D, device='cuda'); ee = torch.randn(n_rec, D, device='cuda'); rel = torch.randn(T, D, device='cuda')
out = torch.empty(n_rec // 4, D, device='cuda'); carry = torch.empty(1, D, device='cuda')
row = np.repeat(np.arange(n_rec // 4, dtype=np.uint32), 4)
flags = row.copy(); flags[0::4] |= np.uint32(1 << 30); flags[3::4] |= np.uint32(1 << 31)
rowflags = torch.from_numpy(flags.view(np.int32)).cuda()
chunks = torch.full((n_rec // 32, 2), -1, dtype=torch.int32, device='cuda')
for eid_name, eid in (('sequential', np.arange(n_rec)), ('random', rng.permutation(n_rec))):
    for src_name, src in (('sequential', np.arange(n_rec) % N), ('random', rng.integers(0, N, n_rec))):
        rec = np.zeros((n_rec, 4), dtype=np.int32)
        rec[:, 0] = eid; rec[:, 1] = src; rec[:, 2] = rng.integers(0, T, n_rec); rec[:, 3] = np.float32(0.5).view(np.int32)
        recd = torch.from_numpy(rec).cuda()
        fn = lambda: L.call('kgc_agg_fwd', p(x), p(rel), T, p(ee), p(recd), p(rowflags), p(chunks), n_rec, p(out), p(carry), D, st())
        fn(); torch.cuda.synchronize()
        us = timed(fn)
        by = n_rec * (400 + 20) + N * 400 + n_rec // 4 * 400
        print('edge ids %-10s source rows %-10s : %6.1f us  %5.0f GB/s algorithmic' % (eid_name, src_name, us, by / us / 1e3))
