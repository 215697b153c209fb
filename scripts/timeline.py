"""Diagnostic: per-step timeline of the fused forward kernel (IBC_DBG_TIMELINE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['IBC_DBG_TIMELINE'] = '1'
import torch
from ibc_b200 import engine as eng
e = eng.EBMEngine(20, 2, 128, 16)
p = (torch.randn(e.param_count) * 0.05).cuda(); e.bind_params(p)
rows = 131584
obs = torch.randn(512, 20).cuda(); act = torch.rand(rows, 2).cuda()
for _ in range(2):
  e.energy(obs, act, 257)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == 'train':
  # the operand-saving instance of the forward (training step)
  e.train_grads(obs, act, 256, eng.loss_cfg(add_grad_penalty=False), 512)
  torch.cuda.synchronize()
