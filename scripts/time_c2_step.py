"""C2 training gradients (B=512, N=256, W=128, D=16) timed per phase with the handle's
CUDA events: block-major fused backward (default) against round 1's split kernels
(IBC_TRAIN_SPLIT=1).  `python scripts/time_c2_step.py [steps]`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ibc_b200 import engine as eng


def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
  dev = torch.device('cuda', 0)
  torch.cuda.set_device(0)
  e = eng.EBMEngine(20, 2, 128, 16, dev)
  params = (torch.randn(e.param_count, device=dev) * 0.05).contiguous()
  e.bind_params(params)
  b, n = 512, 256
  obs = torch.randn(b, 20, device=dev)
  acts = torch.rand(b * (n + 1), 2, device=dev) * 2.2 - 1.1
  cfg = eng.loss_cfg(1.0, False)
  for mode in ('fused', 'split'):
    if mode == 'split':
      os.environ['IBC_TRAIN_SPLIT'] = '1'
    else:
      os.environ.pop('IBC_TRAIN_SPLIT', None)
    for _ in range(3):
      e.train_grads(obs, acts, n, cfg, b)
    torch.cuda.synchronize()
    e.set_profiling(True)
    acc = np.zeros(5)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
      e.train_grads(obs, acts, n, cfg, b)
      acc += np.array(e.train_timings())
    t1.record()
    torch.cuda.synchronize()
    e.set_profiling(False)
    print(mode, 'status', e.kernel_status(), 'ms/call %.4f' % (t0.elapsed_time(t1) / steps),
          'phases fwd/nce/bwd/wgrad/first-last', ['%.4f' % x for x in acc / steps], flush=True)


if __name__ == '__main__':
  main()
