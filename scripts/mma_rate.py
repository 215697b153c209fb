"""Diagnostic: tcgen05.mma issue rate by operand source / layout (cycles per 128xNx8 tf32 MMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
for n in (128,):
  for mode, name in ((20, 'SS no-swizzle'), (21, 'TS (A in TMEM)'), (22, 'SS 128B-swizzle descriptors'), (25, 'SS two N-half accumulators'), (26, 'SS four N-quarter accumulators'), (27, 'SS two alternating full accumulators'), (30, 'clean SS'), (31, 'clean TS'), (32, 'clean SS 2 acc'), (33, 'clean TS 2 acc'), (34, 'SS + smem-store hammer'), (35, 'TS + tmem ld/st hammer'), (36, 'SS + tmem ld/st hammer'), (37, 'TS + smem-store hammer'), (38, 'SS commit every 4'), (39, 'SS commit every 16'), (40, 'SS + bulk-copy flood'), (41, 'TS + bulk-copy flood'), (42, 'SS + 1x16KB bulk'), (43, 'SS + 1x64KB bulk'), (50, 'clean SS, all SMs'), (51, 'clean TS, all SMs'), (55, 'TS + tmem hammer, all SMs'), (61, 'TS + bulk flood, all SMs')):
    if mode < 50 and mode not in (30, 31): continue
    if mode < 30: continue
    if mode == 33 and n > 128: continue
    if mode >= 34 and n != 128: continue
    a = torch.randn(128, 128).cuda(); bt = torch.randn(n, 128).cuda()
    res = [eng.umma_selftest(a, bt, mode)[1] / 10.0 for _ in range(4)]
    print('N=%3d %-28s cycles/MMA %s (ideal %d)' % (n, name, res, n // 2), flush=True)
