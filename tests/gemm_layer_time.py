"""Debug aid (not a test): the layer's dense transforms alone at a BASELINE shape (K4b forward / backward batches, K4c batch),
CUDA events, L2 flushed.   python tests/gemm_layer_time.py [wikidata5m]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import kgc_gcn_b200 as k
from kgc_gcn_b200 import conv as C
dev = torch.device('cuda', 0)
N = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'wikidata5m'][0]
D, Do = 100, 200
flush = bench.Flusher(dev)
a3 = [torch.randn(N, D, device=dev) * 0.05 for _ in range(3)]
d3 = [torch.randn(N, Do, device=dev) * 0.05 for _ in range(3)]
w = [torch.randn(D, Do, device=dev) * 0.1 for _ in range(3)]
pf = [C._pack_b(x) for x in w]
pb = [C._pack_b(x.t()) for x in w]
r3 = [torch.empty(N, Do, device=dev) for _ in range(3)]
g3 = [torch.empty(N, D, device=dev) for _ in range(3)]
dw = [torch.empty(D, Do, device=dev) for _ in range(3)]
print('K4b fwd  %.3f ms' % bench.time_kernel(lambda: C.gemm_nt_batch(a3, pf, r3), flush, iters=5))
print('K4b bwd  %.3f ms' % bench.time_kernel(lambda: C.gemm_nt_batch(d3, pb, g3), flush, iters=5))
print('K4c      %.3f ms' % bench.time_kernel(lambda: C.gemm_tn_batch(a3, d3, dw), flush, iters=5))
