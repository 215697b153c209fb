"""Diagnostic: two 4-iteration Langevin chains on the PixelEBM value head (1024 rows) for the
ncu launch list / --set full captures under profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
dev = torch.device('cuda', 0)
pe = eng.PixelEngine(2, 240, 320, 3, 180, 240, 2, 1024, 1, dev)
pe.bind_params((torch.randn(pe.param_count, device=dev) * 0.05).contiguous())
mn, mx = torch.full((2,), -1.1, device=dev), torch.full((2,), 1.1, device=dev)
cfg = eng.langevin_cfg(num_iterations=4)
enc = torch.rand(128, 256, device=dev)
acts = (torch.rand(1024, 2, device=dev) * 2.2 - 1.1)
pe.langevin(enc, acts, 8, mn, mx, cfg, None, 1)
pe.langevin(enc, acts, 8, mn, mx, cfg, None, 2)
torch.cuda.synchronize()
