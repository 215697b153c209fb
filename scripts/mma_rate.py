"""Diagnostic: tcgen05.mma issue rate by operand source / layout (cycles per 128xNx8 tf32 MMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
for n in (128,):
  for mode, name in ((20, 'SS no-swizzle'), (21, 'TS (A in TMEM)'), (22, 'SS 128B-swizzle descriptors'), (23, 'TS A at col 128 (D at 0)'), (24, 'TS A at col 384 (D at 0)')):
    a = torch.randn(128, 128).cuda(); bt = torch.randn(n, 128).cuda()
    d, st = eng.umma_selftest(a, bt, mode)
    print('N=%3d %-28s cycles/MMA %d (ideal %d)' % (n, name, st, n // 2), flush=True)
