"""Small profiling driver (run under ncu by gpurun): a few train steps, one
stand-alone fused energy forward on the step's rows, one DFO policy action."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

def main():
  steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
  torch.cuda.set_device(0)
  device = torch.device('cuda', 0)
  c = bench.CFG
  net, agent, policy = bench.build_agent(c, device)
  batch = bench.synthetic_batch(c, c['batch'], 100, device=device)
  for _ in range(steps):
    info = agent.train((batch, ()))
  torch.cuda.synchronize()
  rows = c['batch'] * (c['num_counter_examples'] + 1)
  obs_flat = net._flat_obs(batch[0])
  acts = torch.rand(rows, c['act_dim'], device=device) * 2.2 - 1.1
  e = net.engine.energy(obs_flat, acts, c['num_counter_examples'] + 1)
  torch.cuda.synchronize()
  from ibc_b200.ibc import specs
  ts = specs.restart({k: v[:1] for k, v in batch[0].items()})
  a = policy.action(ts)
  torch.cuda.synchronize()
  print('loss', float(info.loss), 'E0', float(e[0]), 'action', a.action.cpu().tolist(),
        'status', net.engine.kernel_status())

if __name__ == '__main__':
  main()
