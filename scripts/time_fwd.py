"""Energy forward (W=128, D=16) timed with CUDA events on the C2 step's 131 584 rows and on one
DFO pass's 16 384 rows.  IBC_FWD_ROWS=1 selects the one-thread-per-row kernel for comparison."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
e = eng.EBMEngine(20, 2, 128, 16, dev)
e.bind_params((torch.randn(e.param_count, device=dev) * 0.05).contiguous())
for b, n in ((512, 257), (1, 16384), (8, 16384)):
  obs = torch.randn(b, 20, device=dev)
  act = torch.rand(b * n, 2, device=dev) * 2.2 - 1.1
  for _ in range(5):
    en = e.energy(obs, act, n)
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  reps = 50
  t0.record()
  for _ in range(reps):
    en = e.energy(obs, act, n)
  t1.record()
  torch.cuda.synchronize()
  ms = t0.elapsed_time(t1) / reps
  print('rows %7d  %.4f ms  %.1f M rows/s  E0 %.6f sum %.4f status %d' % (
      b * n, ms, b * n / ms / 1e3, float(en[0]), float(en.double().sum()), e.kernel_status()), flush=True)
