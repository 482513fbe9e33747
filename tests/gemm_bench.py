"""Debug aid (not a test): cold-L2 CUDA-event timing of the dense transforms at the WN18RR layer shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgc_gcn_b200 as k

def timed(fn, reps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for _ in range(reps):
        flush.zero_(); flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

M = 40943
for K, N in ((100, 200), (200, 100)):
    a = torch.randn(M, K, device='cuda'); b = torch.randn(K, N, device='cuda'); out = torch.empty(M, N, device='cuda')
    for _ in range(3): k.gemm_nt(a, b, out)
    ref = a.double() @ b.double()
    err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    us = timed(lambda: k.gemm_nt(a, b, out))
    gb = M * (K + N) * 4 / 1e9
    print('gemm_nt  M=%d K=%d N=%d: %.1f us  (%.0f GB/s algorithmic)  err %.2e' % (M, K, N, us, gb / us * 1e6, err))
for Ka, Nb in ((100, 200),):
    a = torch.randn(M, Ka, device='cuda'); b = torch.randn(M, Nb, device='cuda'); out = torch.empty(Ka, Nb, device='cuda')
    for tc in (True, False):
        for _ in range(3): k.gemm_tn(a, b, out, tensor_cores=tc)
        us = timed(lambda: k.gemm_tn(a, b, out, tensor_cores=tc))
        print('gemm_tn tc=%s M=%d Ka=%d Nb=%d: %.1f us (%.0f GB/s algorithmic)' % (tc, M, Ka, Nb, us, M * (Ka + Nb) * 4 / 1e3 / us))
