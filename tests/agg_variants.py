"""Debug aid (not a test): time the three level-0 aggregation kernels of one graph shape through the C ABI and print a
digest of their outputs.  profiles/r01_agg_variants.md was produced with it from probe builds of agg.cu that selected a
kernel variant through KGC_AGG_VARIANT / KGC_AGG_KU / KGC_AGG_MINB (the variables are ignored by the product build, which
always runs agg_lean_kernel); the digest is how "bit-identical to the baseline" was checked for every variant."""
import hashlib, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import mgcn_oracle as orc
import kgc_gcn_b200 as k
L = k._lib
p, st = L.ptr, L.stream
shape = sys.argv[1] if len(sys.argv) > 1 else 'wn'
N, R, E, seed = {'wn': (40943, 11, 86835, 0), 'fb': (14541, 237, 272115, 1)}[shape]
D, T = 100, 2 * R + 1
tri = orc.synthetic_triples(N, R, E, seed)
g = orc.build_graph(tri, N, R)
dev = torch.device('cuda')
ei, et = torch.from_numpy(g['edge_index']).to(dev), torch.from_numpy(g['edge_attr'][0]).to(dev)
gen = torch.Generator().manual_seed(0)
x = torch.randn(N, D, generator=gen).to(dev); ee = torch.randn(2 * E, D, generator=gen).to(dev)
relp = torch.randn(T, D, generator=gen).to(dev); g3 = torch.randn(3, N, D, generator=gen).to(dev)
plan = k.get_plan(ei, et, N, T)
agg = torch.zeros((2, N, D), device=dev); d_ee = torch.zeros_like(ee); d_x = torch.zeros((N, D), device=dev)
d_rel = torch.zeros((T, D), device=dev)
sf, ss, sr = plan.fwd, plan.bwd_src, plan.bwd_rel
pf, ps, pr = (torch.zeros((max(s.n_carry, 1), D), device=dev) for s in (sf, ss, sr))
fns = {
    'fwd': lambda: L.call('kgc_agg_fwd', p(x), p(relp), T, p(ee), p(plan.rec_dst), p(sf.rowflags), p(sf.chunks), sf.n_rec,
                          p(agg), p(pf), D, st()),
    'bwd_src': lambda: L.call('kgc_agg_bwd_src', p(x), p(relp), T, p(ee), p(g3), p(plan.rec_src), p(ss.rowflags), p(ss.chunks),
                              ss.n_rec, N, E, p(g3[2]), p(d_ee), p(d_x), p(ps), D, st()),
    'bwd_rel': lambda: L.call('kgc_agg_bwd_rel', p(x), p(ee), p(g3), p(plan.rec_type), p(sr.rowflags), p(sr.chunks), sr.n_rec,
                              N, E, p(d_rel), p(pr), D, st()),
}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev); fsrc = torch.zeros(64 << 20, device=dev)
out = {'variant': os.environ.get('KGC_AGG_VARIANT', '0'), 'ctas': os.environ.get('KGC_AGG_CTAS', ''), 'shape': shape}
for name, fn in fns.items():
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(int(os.environ.get('REPS', '15'))):
        flush.zero_(); fsrc.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    out[name + '_us'] = round(ts[len(ts) // 2], 2); out[name + '_min'] = round(ts[0], 2)
h = hashlib.sha1()
for t in (agg, pf, d_ee, d_x, ps, d_rel, pr):
    h.update(t.cpu().numpy().tobytes())
out['digest'] = h.hexdigest()[:12]
print(json.dumps(out))
