"""Debug aid (not a test): per-role clock64 timeline of CTA 0 of the tensor-core weight-gradient kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgc_gcn_b200 as k
L = k._lib
M, Ka, Nb = 40943, 100, 200
a = torch.randn(M, Ka, device='cuda'); b = torch.randn(M, Nb, device='cuda'); out = torch.empty(Ka, Nb, device='cuda')
for _ in range(3): k.gemm_tn(a, b, out)
dbg = torch.zeros(9 * 64, dtype=torch.int64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); flush.zero_(); flush.sum()
dbg[8 * 64] = 1 << 62
L.lib().kgc_gemm_set_debug(L.ptr(dbg))
k.gemm_tn(a, b, out)
torch.cuda.synchronize()
L.lib().kgc_gemm_set_debug(None)
env = dbg.cpu().view(9, 64)[8]
d = dbg.cpu().view(9, 64)[:8]
t0 = int(d[d > 0].min())
print('all-CTA envelope %.1f us; CTA0 entry->exit %.1f us = %d clk; CTA0 entry->first TMA %d clk' % (
    (int(env[1]) - int(env[0])) / 1e3, (int(env[4]) - int(env[2])) / 1e3, int(env[5]) - int(env[3]), t0 - int(env[3])))
names = ['tma_issue', 'split_raw_ready', 'split_lo_free', 'split_done', 'mma_split_ready', 'mma_issued', 'epi_acc_ready', 'epi_done']
for r, n in enumerate(names):
    v = [int(x) - t0 for x in d[r] if int(x) > 0]
    print('%-16s' % n, v[:24])
