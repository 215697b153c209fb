"""Diagnostic: one C5 PixelEBM train step (pixel_ebm_best.gin shapes) for an ncu launch
list: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python
scripts/profile_pixel.py [batch] [fp32]`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import _lib
from ibc_b200 import engine as eng

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
e = eng.PixelEngine(2, 240, 320, 3, 180, 240, 2, 1024, 1)
if len(sys.argv) > 2 and sys.argv[2] == 'fp32':
  e.set_precision(_lib.PRECISION_FP32)
flat = (torch.randn(e.param_count) * 0.05).cuda()
e.bind_params(flat)
img = torch.randint(0, 256, (b, 2, 240, 320, 3), dtype=torch.uint8).cuda()
act = (torch.rand(b * 257, 2) * 2 - 1).cuda()
cfg = eng.loss_cfg(1.0, False)
for _ in range(2):
  per, _, grads = e.train_grads(img, act, 256, cfg, b)
torch.cuda.synchronize()
print('status', e.kernel_status(), 'loss', float(per.mean()), 'finite', bool(torch.isfinite(grads).all()))
