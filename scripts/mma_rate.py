"""Diagnostic: tcgen05.mma issue rate by operand source / layout (cycles per 128xNx8 tf32 MMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
for n in (128,):
  for mode, name in ((20, 'SS no-swizzle'), (21, 'TS (A in TMEM)'), (22, 'SS 128B-swizzle descriptors'), (25, 'SS two N-half accumulators'), (26, 'SS four N-quarter accumulators'), (27, 'SS two alternating full accumulators')):
    a = torch.randn(128, 128).cuda(); bt = torch.randn(n, 128).cuda()
    res = [eng.umma_selftest(a, bt, mode)[1] / 10.0 for _ in range(4)]
    print('N=%3d %-28s cycles/MMA %s (ideal %d)' % (n, name, res, n // 2), flush=True)
