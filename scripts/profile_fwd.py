"""The inference energy forward (mlp_fwd_tmem_kernel<0>) on the C2 step's 131 584 rows, a few
launches: run under `ncu --set full` for the tensor-pipe counters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng

torch.cuda.set_device(0)
dev = torch.device('cuda', 0)
e = eng.EBMEngine(20, 2, 128, 16, dev)
e.bind_params((torch.randn(e.param_count, device=dev) * 0.05).contiguous())
b, n = 512, 257
obs = torch.randn(b, 20, device=dev)
act = torch.rand(b * n, 2, device=dev) * 2.2 - 1.1
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
  en = e.energy(obs, act, n)
torch.cuda.synchronize()
print('E0', float(en[0]), 'status', e.kernel_status())
