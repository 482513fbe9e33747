"""Debug aid (not a test): which output positions the MN-major tensor-core reduction fills for one-hot inputs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kgc_gcn_b200 as k

M, Ka, Nb = 32, 100, 200
def probe(m_a, i_a, name):
    a = torch.zeros(M, Ka); b = torch.zeros(M, Nb)
    a[m_a, i_a] = 1.0
    b[m_a] = torch.arange(1, Nb + 1).float()
    out = torch.full((Ka, Nb), float('nan'), device='cuda')
    k.gemm_tn(a.cuda(), b.cuda(), out, tensor_cores=True)
    out = out.cpu()
    nz = torch.nonzero(out != 0)
    rows = sorted(set(nz[:, 0].tolist()))
    print(name, 'expect row', i_a, '-> nonzero rows', rows[:12], 'count', nz.shape[0])
    for r in rows[:3]:
        vals = out[r]
        idx = torch.nonzero(vals != 0).flatten().tolist()
        print('   row', r, 'cols', idx[:10], '...', idx[-3:], 'vals', [round(float(vals[c]), 2) for c in idx[:10]])
probe(0, 0, 'A[0][0]')
probe(0, 5, 'A[0][5]')
probe(3, 0, 'A[3][0]')
probe(9, 0, 'A[9][0]')
probe(0, 40, 'A[0][40]')
probe(17, 99, 'A[17][99]')
