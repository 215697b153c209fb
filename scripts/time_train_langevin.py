"""Diagnostic: the loss half of a Langevin-config train step (InfoNCE + gradient penalty
forward / reverse / tangent / weight gradients on B*(N+1) = 4608 rows) per precision."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import _lib
from ibc_b200 import engine as eng

for name, (o, a, w, d) in (('C3 particle-32D W=256 D=2', (256, 32, 256, 2)),
                           ('C3-best W=2048 D=2', (256, 32, 2048, 2)),
                           ('C4 d4rl door W=512 D=8', (39, 28, 512, 8))):
  b, n = 512, 8
  e = eng.EBMEngine(o, a, w, d)
  e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
  obs = torch.randn(b, o).cuda()
  acts = (torch.rand(b * (n + 1), a) * 2.2 - 1.1).cuda()
  cfg = eng.loss_cfg(1.0, True, 'inf', 1.0, True, 1.0)
  for prec, label in ((_lib.PRECISION_TF32, 'tf32'), (_lib.PRECISION_FP32, 'fp32')):
    e.set_precision(prec)
    for _ in range(2):
      e.train_grads(obs, acts, n, cfg, b)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
      out = e.train_grads(obs, acts, n, cfg, b)
    t.record(); torch.cuda.synchronize()
    print('%-28s train_grads + grad penalty %.3f ms [%s] status %d finite %s'
          % (name, s.elapsed_time(t) / 5, label, e.kernel_status(),
             bool(torch.isfinite(out[-1]).all())))
