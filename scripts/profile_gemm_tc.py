"""Diagnostic: one launch of each form of the general tf32 GEMM at 4096^3 (and the conv2
shape) for `ncu --set full -k regex:gemm_tc`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

m = n = k = 4096
a = torch.randn(m, k, device='cuda'); w = torch.randn(k, n, device='cuda')
g = torch.randn(m, n, device='cuda'); dw = torch.zeros(k, n, device='cuda')
for _ in range(2):
  eng.gemm_tf32(a, w, m, n, k, check=False)
  eng.gemm_tf32(g, w, m, k, n, trans_b=True, check=False)
  eng.gemm_tf32(a, g, k, n, m, trans_a=True, out=dw, check=False)
torch.cuda.synchronize()
print('ok')
