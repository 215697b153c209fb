"""Diagnostic: Langevin chain time with and without CUDA-graph replay (run twice:
plain and with IBC_NO_GRAPHS=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

for name, (o, a, w, d, b, n) in (('pushing_states W=128 D=16 R=4096', (20, 2, 128, 16, 512, 8)),
                                 ('C3 W=256 D=2 R=4096', (256, 32, 256, 2, 512, 8)),
                                 ('particle inference W=256 D=2 R=512', (16, 2, 256, 2, 1, 512))):
  e = eng.EBMEngine(o, a, w, d)
  e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
  obs = torch.randn(b, o).cuda()
  mn, mx = torch.full((a,), -1.1).cuda(), torch.full((a,), 1.1).cuda()
  cfg = eng.langevin_cfg(num_iterations=100)
  acts = (torch.rand(b * n, a) * 2.2 - 1.1).cuda()
  for i in range(3):
    e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=i)
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
  s.record()
  for i in range(10):
    e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=10 + i)
  t.record(); torch.cuda.synchronize()
  print('%-40s %.3f ms/chain (%s)' % (name, s.elapsed_time(t) / 10,
        'no graphs' if os.environ.get('IBC_NO_GRAPHS') else 'graph replay'))

# wide stacks (per-layer tensor-core GEMMs, mlp_wide.cu) against the fp32 FMA path
from ibc_b200 import _lib
for name, (o, a, w, d, b, n) in (('C4 d4rl W=512 D=8 R=4096', (39, 28, 512, 8, 512, 8)),
                                 ('particle best W=2048 D=2 R=4096', (16, 2, 2048, 2, 512, 8))):
  e = eng.EBMEngine(o, a, w, d)
  e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
  obs = torch.randn(b, o).cuda()
  acts = (torch.rand(b * n, a) * 2.2 - 1.1).cuda()
  for prec, label in ((_lib.PRECISION_TF32, 'tf32 tcgen05'), (_lib.PRECISION_FP32, 'fp32 FMA')):
    e.set_precision(prec)
    for i in range(2):
      e.energy_dact(obs, acts, n)
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(5):
      e.energy_dact(obs, acts, n)
    t.record(); torch.cuda.synchronize()
    print('%-36s energy + dE/dact %.3f ms (%s)' % (name, s.elapsed_time(t) / 5, label))

# C4 chain (wide path, graph replay)
o, a, w, d, b, n = 39, 28, 512, 8, 512, 8
e = eng.EBMEngine(o, a, w, d)
e.bind_params((torch.randn(e.param_count) * 0.05).cuda())
obs = torch.randn(b, o).cuda()
mn, mx = torch.full((a,), -1.1).cuda(), torch.full((a,), 1.1).cuda()
cfg = eng.langevin_cfg(num_iterations=100)
acts = (torch.rand(b * n, a) * 2.2 - 1.1).cuda()
for i in range(4):
  e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=i)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
s.record()
for i in range(5):
  e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=10 + i)
t.record(); torch.cuda.synchronize()
print('C4 d4rl W=512 D=8 R=4096 chain        %.3f ms/chain (%s)' % (s.elapsed_time(t) / 5,
      'no graphs' if os.environ.get('IBC_NO_GRAPHS') else 'graph replay'))
