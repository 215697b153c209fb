"""Small profiling driver for the sampler kernels (run under ncu by gpurun): one Langevin chain
of the pushing_states net (langevin_f16_kernel, W = 128, D = 16, 4096 rows), one of the C3 net
(langevin256_kernel, W = 256, D = 2, 4096 rows) and one DFO action (resample_cluster_kernel,
16 384 candidates).  K is kept short: ncu replays every kernel ~40 times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

k = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.cuda.set_device(0)
for name, (o, a, w, d, b, n) in (('pushing_states', (20, 2, 128, 16, 512, 8)),
                                 ('C3', (256, 32, 256, 2, 512, 8))):
  e = eng.EBMEngine(o, a, w, d)
  e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
  obs = torch.randn(b, o).cuda()
  mn, mx = torch.full((a,), -1.1).cuda(), torch.full((a,), 1.1).cuda()
  acts = (torch.rand(b * n, a) * 2.2 - 1.1).cuda()
  e.langevin(obs, acts, n, mn, mx, eng.langevin_cfg(num_iterations=k), seed=3)
  torch.cuda.synchronize()
  print(name, 'chain status', e.kernel_status(), float(acts.abs().mean()))
e = eng.EBMEngine(20, 2, 128, 16)
e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
obs = torch.randn(1, 20).cuda()
mn, mx = torch.full((2,), -1.1).cuda(), torch.full((2,), 1.1).cuda()
acts = (torch.rand(16384, 2) * 2.2 - 1.1).cuda()
e.dfo(obs, acts, 16384, mn, mx, seed=5)
torch.cuda.synchronize()
print('dfo status', e.kernel_status(), float(acts.abs().mean()))
