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

# K-major operands (the chain GEMMs): D = A[128,K] . B[N,K]^T; A from shared memory (100) or from
# tensor memory, two K-elements per 32-bit column: even k in the low half (101) / odd k low (102)
print('K-major probes: mode 100 = SS, 101 = TS (even k low), 102 = TS (odd k low)')
for k, n in ((16, 64), (32, 128), (128, 128)):
  a = torch.randn(128, k, generator=g)
  b = torch.randn(n, k, generator=g)
  want = a.half().double() @ b.half().double().t()
  for mode in (100, 101, 102):
    # the engine wrapper takes A as [K,128] / B as [K,N] shaped buffers: only the byte count matters
    d, st = eng.umma_f16_selftest(a.reshape(k, 128).cuda(), b.reshape(k, n).cuda(), layout_type=mode)
    err = (d.double().cpu() - want).abs().max().item()
    print('K=%d N=%d mode %d status %d max err %.3g (|want| max %.3g)' % (k, n, mode, st, err, want.abs().max().item()), flush=True)

# MN-major A WITHOUT swizzle: the forward's saved-tile layout [chunk of 8][rows][16 B] as is
print('MN-major no-swizzle A probes: 200 = LBO 128 / SBO 16 K, 201 = swapped')
for k, n in ((32, 128), (128, 128)):
  a = torch.randn(k, 128, generator=g)
  b = torch.randn(k, n, generator=g)
  want = a.half().double().t() @ b.half().double()
  for mode in (200, 201):
    d, st = eng.umma_f16_selftest(a.cuda(), b.cuda(), layout_type=mode)
    err = (d.double().cpu() - want).abs().max().item()
    print('K=%d N=%d mode %d status %d max err %.3g (|want| max %.3g)' % (k, n, mode, st, err, want.abs().max().item()), flush=True)
