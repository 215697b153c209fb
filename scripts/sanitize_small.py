"""Small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck` (slow under
the tool: tiny shapes): fused train step (kind::f16 forward + block-major backward), DFO with the
cluster resample / softmax, the kind::f16 Langevin chain, the one-launch helpers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
e = eng.EBMEngine(20, 2, 128, 4, dev)
e.bind_params((torch.randn(e.param_count, device=dev) * 0.05).contiguous())
b, n = 6, 50                                     # 306 rows: two full tiles and a ragged one
obs = torch.randn(b, 20, device=dev)
mn = torch.full((2,), -1.1, device=dev); mx = torch.full((2,), 1.1, device=dev)
act = eng.counter_examples_uniform(torch.rand(b, 2, device=dev) * 2 - 1, n, mn, mx, seed=3)
per, _, en, g = e.train_grads(obs, act, n, eng.loss_cfg(1.0, False), b, want_energies=True)
print('train_grads', float(per.sum()), float(g.abs().sum()), e.kernel_status())
a = (torch.rand(2 * 4096, 2, device=dev) * 2.2 - 1.1)
p = e.dfo(obs[:2].contiguous(), a, 4096, mn, mx, seed=5)
print('dfo', float(p.sum()), float(eng.greedy_actions(p.reshape(2, -1), a.reshape(2, 4096, 2), mn, mx).abs().sum()), e.kernel_status())
a = (torch.rand(b * 30, 2, device=dev) * 2.2 - 1.1)
e.langevin(obs, a, 30, mn, mx, eng.langevin_cfg(num_iterations=3), seed=7)
print('langevin', float(a.abs().sum()), e.kernel_status())
torch.cuda.synchronize()
