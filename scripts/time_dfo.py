"""One DFO policy action (B observations x 16 384 candidates, 3 iterations) through ibc_dfo,
timed with CUDA events: C2 inference (W=128, D=16)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
e = eng.EBMEngine(20, 2, 128, 16, dev)
e.bind_params((torch.randn(e.param_count, device=dev) * 0.05).contiguous())
mn = torch.full((2,), -1.1, device=dev); mx = torch.full((2,), 1.1, device=dev)
for b in (1, 8):
  n = 16384
  obs = torch.randn(b, 20, device=dev)
  act0 = torch.rand(b * n, 2, device=dev) * 2.2 - 1.1
  act = act0.clone()
  for i in range(5):
    act.copy_(act0); e.dfo(obs, act, n, mn, mx, seed=i)
  torch.cuda.synchronize()
  reps = 100
  t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0.record()
  for i in range(reps):
    e.dfo(obs, act, n, mn, mx, seed=100 + i)
  t1.record(); torch.cuda.synchronize()
  print('B=%d  ibc_dfo %.1f us per action batch, status %d' % (b, t0.elapsed_time(t1) / reps * 1e3, e.kernel_status()))
