"""Hardware probe of the kind::f16 MN-major descriptor (mlp_bwd_wgrad.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

torch.cuda.set_device(0)
g = torch.Generator().manual_seed(0)
for k, n in ((32, 128), (16, 64), (64, 128), (128, 256)):
  a = torch.randn(k, 128, generator=g)
  b = torch.randn(k, n, generator=g)
  want = a.half().double().t() @ b.half().double()
  for name, kw in (('default', {}),):
    d, st = eng.umma_f16_selftest(a.cuda(), b.cuda(), **kw)
    err = (d.double().cpu() - want).abs().max().item()
    print('K=%d N=%d %-10s status %d max err %.3g (|want| max %.3g)' % (k, n, name, st, err, want.abs().max().item()), flush=True)
