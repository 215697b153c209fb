"""Diagnostic: CUDA-event times of the pieces of one DFO policy action (B=1, N=16384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

def timed(fn, reps=50):
  for _ in range(5): fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(reps): fn()
  e.record(); torch.cuda.synchronize()
  return s.elapsed_time(e) / reps * 1e3

n = 16384
e = eng.EBMEngine(20, 2, 128, 16)
p = (torch.randn(e.param_count) * 0.05).cuda(); e.bind_params(p)
obs = torch.randn(1, 20).cuda(); act = torch.rand(n, 2).cuda() * 2 - 1
logits = e.energy(obs, act, n).reshape(1, n).contiguous()
u = torch.rand(1, n, dtype=torch.float64).cuda()
print('energy forward  %7.1f us' % timed(lambda: e.energy(obs, act, n)))
print('resample        %7.1f us' % timed(lambda: eng.resample(logits, u)))
print('resample (peaked logits x50) %7.1f us' % timed(lambda: eng.resample((logits * 50).contiguous(), u)))
print('softmax_rows    %7.1f us' % timed(lambda: eng.softmax_rows(logits, 1.0)))
rows, _, _ = eng.resample(logits, u)
mn = torch.full((2,), -1.1).cuda(); mx = torch.full((2,), 1.1).cuda()
print('gather_perturb  %7.1f us' % timed(lambda: eng.gather_perturb(act, rows, 0.33, mn, mx, None, 1)))
