"""Diagnostic: per-layer cycle timeline of langevin_f16_kernel (CTA 0, second iteration) on the
pushing_states net: run with IBC_DBG_TIMELINE=1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
o, a, w, d, b, n = 20, 2, 128, 16, 512, 8
e = eng.EBMEngine(o, a, w, d)
e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
obs = torch.randn(b, o).cuda()
mn, mx = torch.full((a,), -1.1).cuda(), torch.full((a,), 1.1).cuda()
cfg = eng.langevin_cfg(num_iterations=4)
acts = (torch.rand(b * n, a) * 2.2 - 1.1).cuda()
e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=1)
torch.cuda.synchronize()
