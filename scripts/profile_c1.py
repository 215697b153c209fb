"""Diagnostic: launch list of the C1 (particle 2-D, W=256 D=2, B=512 N=256) gradient pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
e = eng.EBMEngine(16, 2, 256, 2)
e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
b, n = 512, 256
obs = torch.randn(b, 16).cuda()
acts = (torch.rand(b * (n + 1), 2) * 2 - 1).cuda()
cfg = eng.loss_cfg(1.0, False)
for _ in range(3):
  e.train_grads(obs, acts, n, cfg, b)
torch.cuda.synchronize()
