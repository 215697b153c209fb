"""Diagnostic: pushing_pixels/pixel_ebm_langevin.gin shapes on the late-fusion PixelEBM --
value-head energy + dE/dact, the 100-iteration Langevin chain (training rows 128 x 8,
inference rows 1 x 2048), the train step with gradient penalty, and a policy action."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import _lib
from ibc_b200 import engine as eng


def timed(fn, reps):
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
  s.record()
  for i in range(reps):
    fn(i)
  t.record(); torch.cuda.synchronize()
  return s.elapsed_time(t) / reps


dev = torch.device('cuda', 0)
pe = eng.PixelEngine(2, 240, 320, 3, 180, 240, 2, 1024, 1, dev)
pe.bind_params((torch.randn(pe.param_count, device=dev) * 0.05).contiguous())
mn, mx = torch.full((2,), -1.1, device=dev), torch.full((2,), 1.1, device=dev)
cfg = eng.langevin_cfg(num_iterations=100)
for prec, label in ((_lib.PRECISION_TF32, 'tf32'), (_lib.PRECISION_FP32, 'fp32')):
  pe.set_precision(prec)
  for b, n in ((128, 8), (1, 2048)):
    enc = torch.rand(b, 256, device=dev)
    acts = (torch.rand(b * n, 2, device=dev) * 2.2 - 1.1)
    pe.energy_dact(enc, acts, n)
    t = timed(lambda i: pe.energy_dact(enc, acts, n), 20)
    print('%s  B=%d n=%d  energy + dE/dact %.3f ms' % (label, b, n, t))
    pe.langevin(enc, acts.clone(), n, mn, mx, cfg, None, 1)
    t = timed(lambda i: pe.langevin(enc, acts.clone(), n, mn, mx, cfg, None, 2 + i), 5)
    print('%s  B=%d n=%d  Langevin K=100 %.3f ms/chain (%s)' % (
        label, b, n, t, 'no graphs' if os.environ.get('IBC_NO_GRAPHS') else 'graphs allowed'))
pe.set_precision(_lib.PRECISION_TF32)
b, n = 128, 8
img = torch.randint(0, 256, (b, 2, 240, 320, 3), device=dev, dtype=torch.uint8)
acts = torch.rand(b * (n + 1), 2, device=dev) * 2 - 1
for gp in (False, True):
  c = eng.loss_cfg(1.0, gp)
  pe.train_grads(img, acts, n, c, b)
  t = timed(lambda i: pe.train_grads(img, acts, n, c, b), 5)
  print('train_grads B=128 N=8 grad penalty %s: %.3f ms' % (gp, t))
t = timed(lambda i: pe.encode(img), 5)
print('encode 128 frame pairs: %.3f ms' % t)
