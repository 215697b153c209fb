"""world_size-2 gloo test (CPU) of the data-parallel path: per-rank gradients of
the half batches, scaled by the GLOBAL batch, summed by
ibc_b200.parallel.allreduce_gradients, equal the single-process gradient of the
concatenated batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _grads(flat, spec, obs, act, n, global_batch):
  from oracle import losses as olosses
  from oracle import mlp_ebm as omlp
  theta = flat.clone().requires_grad_(True)
  b = obs.shape[0]
  obs_t = torch.repeat_interleave(obs, n + 1, 0)
  pred = omlp.energy(theta, spec, obs_t, act).reshape(b, n + 1)
  per = olosses.info_nce(pred, b, n, 1.0)
  (g,) = torch.autograd.grad(per.sum() / global_batch, theta)
  return g


def _worker(rank, world, port, out):
  sys.path.insert(0, ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  from ibc_b200 import parallel
  from oracle import mlp_ebm as omlp
  spec = omlp.MLPSpec(8, 2, 32, 2)
  flat = omlp.init_params(spec, seed=1).double()
  g = torch.Generator().manual_seed(0)
  b, n = 8, 7
  obs = torch.randn(b, spec.obs_dim, generator=g).double()
  act = (torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1).double()
  assert parallel.num_replicas() == world and parallel.replica_rank() == rank
  per = parallel.per_replica_batch_size(b)
  lo, hi = rank * per, (rank + 1) * per
  assert parallel.shard_observations(b) == (lo, hi)
  local = _grads(flat, spec, obs[lo:hi], act[lo * (n + 1):hi * (n + 1)], n, b)
  parallel.allreduce_gradients(local)
  full = _grads(flat, spec, obs, act, n, b)
  err = float((local - full).abs().max() / full.abs().max())
  # the peer-memory reducer is for NCCL ranks on CUDA devices: under gloo / on the CPU `create`
  # answers None on every rank (no collective is left half entered) and the agent keeps
  # allreduce_gradients + ibc_adam_step
  assert parallel.PeerReducer.create(flat.numel(), torch.device('cpu')) is None
  if rank == 0:
    with open(out, 'w') as f:
      f.write(repr(err))
  dist.barrier()
  dist.destroy_process_group()


def test_dp_allreduce_matches_single_process(tmp_path):
  out = str(tmp_path / 'err.txt')
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
  assert float(open(out).read()) < 1e-12


def test_parallel_helpers_single_process():
  from ibc_b200 import parallel
  assert parallel.num_replicas() == 1 and parallel.replica_rank() == 0
  t = torch.ones(3)
  assert parallel.allreduce_gradients(t) is t
  assert parallel.shard_observations(10) == (0, 10)
  assert parallel.PeerReducer.create(100, torch.device('cpu')) is None      # single replica
