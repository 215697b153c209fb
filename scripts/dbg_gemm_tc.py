"""Diagnostic: gemm_tc rate with the loaders' copies and / or the MMAs switched off
(IBC_TC_DBG = 1 / 2 / 3, results are garbage): which role bounds a K-slice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

def timed(fn, reps=10, warm=3):
  for _ in range(warm): fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(reps): fn()
  t.record(); torch.cuda.synchronize()
  return s.elapsed_time(t) / reps

for name, m, n, k in (('dense 4096', 4096, 4096, 4096), ('N=32 K=576', 1382400, 32, 576),
                      ('N=64 K=288', 1382400, 64, 288), ('N=32 K=64', 5529600, 32, 64)):
  a = torch.randn(m, k, device='cuda'); w = torch.randn(k, n, device='cuda')
  t = timed(lambda: eng.gemm_tf32(a, w, m, n, k, check=False))
  tiles = ((m + 127) // 128) * ((n + 127) // 128 if n > 64 else 1)
  slices = tiles * ((k + 31) // 32)
  print('IBC_TC_DBG=%s %-12s %.3f ms  %.0f cycles per K-slice per SM (1.9 GHz)'
        % (os.environ.get('IBC_TC_DBG', '0'), name, t, t * 1e-3 * 1.9e9 / (slices / 148)))
