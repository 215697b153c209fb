"""Diagnostic: probe the MN-major UMMA operand layout with one-hot operands."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

N, K = 128, 32
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 3
def probe(m0, k0, n0, k1=None):
  a = torch.zeros(128, K); bt = torch.zeros(N, K)
  a[m0, k0] = 1.0; bt[n0, k0 if k1 is None else k1] = 1.0
  d, st = eng.umma_selftest(a.cuda(), bt.cuda(), MODE)
  nz = d.cpu().nonzero().tolist()
  print('A[%d,%d]=1 B[%d,%d]=1 -> status %d nonzero %s vals %s' % (m0, k0, n0, k0 if k1 is None else k1, st, nz[:6], [float(d[i, j]) for i, j in nz[:6]]), flush=True)

for (m0, k0, n0) in [(0, 0, 0), (1, 0, 0), (4, 0, 0), (5, 0, 3), (0, 1, 0), (0, 7, 0), (0, 8, 0), (0, 9, 2), (37, 13, 77), (127, 31, 127)]:
  probe(m0, k0, n0)
# which B k pairs with A k=0 ?
a = torch.zeros(128, K); a[0, 0] = 1.0
bt = torch.zeros(N, K)
for k in range(K): bt[k, k] = 1.0 + k   # B[n=k, k] = 1+k
d, st = eng.umma_selftest(a.cuda(), bt.cuda(), MODE)
print('A[0,0]=1, B[n,n]=1+n : row0 nonzero', [(j, float(d[0, j])) for j in d[0].cpu().nonzero().reshape(-1).tolist()][:8])
a = torch.zeros(128, K)
for k in range(K): a[k, k] = 1.0 + k
bt = torch.zeros(N, K); bt[0, 0] = 1.0
d, st = eng.umma_selftest(a.cuda(), bt.cuda(), MODE)
print('A[m,m]=1+m, B[0,0]=1 : col0 nonzero', [(i, float(d[i, 0])) for i in d[:, 0].cpu().nonzero().reshape(-1).tolist()][:8])

g = torch.Generator().manual_seed(0)
for (n, k) in [(128, 32), (128, 128), (32, 64), (256, 64)]:
  a = torch.randn(128, k, generator=g); bt = torch.randn(n, k, generator=g)
  d, st = eng.umma_selftest(a.cuda(), bt.cuda(), MODE)
  want = a.double() @ bt.double().t()
  err = (d.cpu().double() - want).abs().max()
  print('random N=%d K=%d status %d max err %.4g (rms %.3g)' % (n, k, st, float(err), float(want.pow(2).mean().sqrt())))
