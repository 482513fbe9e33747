"""Debug aid (not a test): per-role clock64 timeline of CTA 0 of the 3xTF32 GEMM."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgc_gcn_b200 as k
L = k._lib
M, K, N = 40943, 100, 200
a = torch.randn(M, K, device='cuda'); b = torch.randn(K, N, device='cuda'); out = torch.empty(M, N, device='cuda')
for _ in range(3): k.gemm_nt(a, b, out)
dbg = torch.zeros(8 * 64, dtype=torch.int64, device='cuda')
L.lib().kgc_gemm_set_debug(L.ptr(dbg))
k.gemm_nt(a, b, out)
torch.cuda.synchronize()
L.lib().kgc_gemm_set_debug(None)
d = dbg.cpu().view(8, 64)
t0 = int(d[d > 0].min())
names = ['tma_issue', 'split_raw_ready', 'split_lo_free', 'split_done', 'mma_split_ready', 'mma_issued', 'epi_acc_ready', 'epi_done']
for r, n in enumerate(names):
    v = [int(x) - t0 for x in d[r] if int(x) > 0]
    print('%-16s' % n, v[:24])
