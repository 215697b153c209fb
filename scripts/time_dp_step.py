"""C2 data-parallel train step, peer-memory all-reduce + Adam (one launch) against NCCL all-reduce +
ibc_adam_step: torchrun --nproc-per-node N scripts/time_dp_step.py.  Device-timed, max over ranks."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
device = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=device)


def barrier():
  if world > 1:
    dist.barrier()


def mx(x):
  if world == 1:
    return x
  t = torch.tensor([x], device=device, dtype=torch.float64)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  return float(t)


c = bench.CFG
steps = 400
for mode in (('peer', 'nccl') if world > 2 else ('peer', 'nccl', 'peer', 'nccl')):
  os.environ['IBC_PEER_ALLREDUCE'] = '1' if mode == 'peer' else '0'
  net, agent, _ = bench.build_agent(c, device)
  dev = [bench.synthetic_batch(c, c['batch'], 100 + rank * 17 + i, device=device) for i in range(4)]
  host = [bench.synthetic_batch(c, c['batch'], 100 + rank * 17 + i, pinned=True) for i in range(4)]
  for i in range(20):
    agent.train((dev[i % 4], ()))
  t = mx(bench.timed(lambda i: agent.train((dev[i % 4], ())), steps, barrier)) / steps
  # host issue time: 200 steps queued without a sync (the launch queue holds them)
  barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(200):
    agent.train((dev[i % 4], ()))
  th = mx(time.perf_counter() - t0) / 200
  torch.cuda.synchronize()
  if rank == 0:
    print('world %d  %-4s host issue %.4f ms/step' % (world, mode, th * 1e3), flush=True)
  barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(steps):
    float(agent.train((host[i % 4], ())).loss)
  torch.cuda.synchronize()
  te = mx(time.perf_counter() - t0) / steps
  used = getattr(agent, '_peer', None) is not None
  if rank == 0:
    print('world %d  %-4s (peer kernel %s): device %.4f ms/step  e2e %.4f ms/step' % (world, mode, used, t * 1e3, te * 1e3),
          flush=True)
  if used:
    agent._peer.check()
    agent._peer.close()
  del net, agent
if world > 1:
  dist.barrier()
  dist.destroy_process_group()
