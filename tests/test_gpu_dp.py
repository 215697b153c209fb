"""Data-parallel parity on the GPU: two NCCL ranks run the CUDA
`ibc_train_grads` on the two halves of a batch with loss scaling by the GLOBAL batch
(common.aggregate_losses, ibc/agents/ibc_agent.py:246-250), all-reduce the flat gradient
(ibc_b200.parallel.allreduce_gradients, the MirroredStrategy sum of
ibc/agents/base_agent.py:50-58), and must reproduce the one-rank gradient of the whole batch:
1e-5 relative in fp32 (reduction order), 2e-3 per tensor in tf32 (the fp16 dY scale of the
block-major backward is taken per rank).  Skipped with fewer than two GPUs; also covers the
observation-sharded policy (row e2): two ranks' actions == the one-rank batched actions, and the
peer-memory all-reduce + Adam kernel (parallel.PeerReducer) against NCCL all-reduce + Adam."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import math, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, %(root)r)
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
from ibc_b200 import _lib, engine as eng, parallel
from oracle import mlp_ebm as omlp
ok = True
for spec, b, n in ((omlp.MLPSpec(20, 2, 128, 16), 64, 256), (omlp.MLPSpec(39, 28, 512, 4), 32, 8)):
  flat = omlp.init_params(spec, seed=1)
  e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
  dev = flat.cuda().contiguous(); e.bind_params(dev)
  g = torch.Generator().manual_seed(0)
  obs = torch.randn(b, spec.obs_dim, generator=g).cuda()
  act = (torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1).cuda()
  per = parallel.per_replica_batch_size(b)
  lo, hi = parallel.shard_observations(b)
  assert (lo, hi) == (rank * per, (rank + 1) * per)
  gp = spec.width != 128
  cfg = eng.loss_cfg(1.0, gp, 'inf', 0.05, True, 1.0)
  for prec, tol in ((_lib.PRECISION_FP32, 1e-5), (_lib.PRECISION_TF32, 2e-3)):
    e.set_precision(prec)
    _, _, _, local = e.train_grads(obs[lo:hi].contiguous(), act[lo * (n + 1):hi * (n + 1)].contiguous(),
                                   n, cfg, b)
    local = local.clone()
    parallel.allreduce_gradients(local)
    _, _, _, full = e.train_grads(obs, act, n, cfg, b)
    assert e.kernel_status() == 0
    grms = float(full.double().pow(2).mean().sqrt())
    for name, off, shape in omlp.param_layout(spec):
      cnt = math.prod(shape)
      if name == 'be' or name == 'b2_%%d' %% (spec.depth // 2 - 1):
        continue      # both are (sum of all dE) x const = 0 up to rounding: summation-order noise
      w = full[off:off + cnt].double()
      rel = float((local[off:off + cnt].double() - w).norm()) / (float(w.norm()) + 1e-2 * grms * math.sqrt(cnt))
      if rel > tol:
        ok = False
        print('MISMATCH', spec, prec, name, rel, flush=True)
# observation-sharded policy: same draws per observation -> same actions as one rank on the batch
import bench
from ibc_b200.ibc import specs
c = dict(bench.CFG, dfo_samples=2048)
torch.manual_seed(5)
net, agent, policy = bench.build_agent(c, torch.device('cuda', rank), seed=3)
b = 6
obs, _ = bench.synthetic_batch(c, b, 900, device=torch.device('cuda', rank))
ts = specs.restart(obs)
gen = torch.Generator().manual_seed(11)
n = c['dfo_samples']
draws_all = {'uniform_u': torch.rand(b * n, c['act_dim'], generator=gen).cuda(),
             'dfo_uniforms': torch.rand(3, b, n, generator=gen, dtype=torch.float64).cuda(),
             'dfo_normals': torch.randn(2, b * n, c['act_dim'], generator=gen).cuda()}
want = policy.action(ts, draws=draws_all).action
lo, hi = parallel.shard_observations(b)
draws_local = {'uniform_u': draws_all['uniform_u'][lo * n:hi * n].contiguous(),
               'dfo_uniforms': draws_all['dfo_uniforms'][:, lo:hi].contiguous(),
               'dfo_normals': draws_all['dfo_normals'][:, lo * n:hi * n].contiguous()}
got, sl = parallel.sharded_policy_action(policy, ts, draws=draws_local)
if sl != (lo, hi) or got.shape != want.shape or not torch.equal(got, want):
  ok = False
  print('SHARDED POLICY MISMATCH', rank, float((got - want).abs().max()), flush=True)
# peer-memory all-reduce + Adam (parallel.PeerReducer, one launch) against NCCL all-reduce +
# ibc_adam_step on the SAME per-rank gradient vectors, four consecutive steps (both parities
# twice): equal up to the fused-multiply-add contraction of two separately compiled kernels
# (a few ulp: 1e-6 relative + absolute); the replicas themselves stay bit-identical (below)
numel = 268_801
red = parallel.PeerReducer.create(numel, torch.device('cuda', rank))
if red is None:
  ok = False
  print('PEER REDUCER NOT CREATED', rank, flush=True)
else:
  gen = torch.Generator(device='cuda').manual_seed(5)
  state = [torch.randn(numel, device='cuda', generator=gen) for _ in range(3)]
  state[2].abs_()
  mine = [x.clone() for x in state]
  ref = [x.clone() for x in state]
  for step in range(4):
    gvec = torch.randn(numel, device='cuda', generator=torch.Generator(device='cuda').manual_seed(100 * step + rank))
    red.buffer(step).copy_(gvec)
    red.reduce_and_adam(step, mine[0], mine[1], mine[2], 1e-3 * (step + 1), 0.9, 0.999, 1e-7)
    summed = gvec.clone()
    dist.all_reduce(summed)
    eng.adam_step(ref[0], summed, ref[1], ref[2], 1e-3 * (step + 1), 0.9, 0.999, 1e-7)
    red.check()
    close = lambda x, y: torch.allclose(x, y, rtol=1e-6, atol=1e-6)
    if not (close(mine[0], ref[0]) and close(mine[1], ref[1]) and close(mine[2], ref[2])
            and float(red.tail[0]) == float(summed[-1])):
      ok = False
      print('PEER ALLREDUCE + ADAM MISMATCH', rank, step, float((mine[0] - ref[0]).abs().max()), flush=True)
  red.close()
# the agent on the peer path: eight data-parallel train steps from the same seed on the same
# per-rank batches, against the NCCL path; the gradients carry atomics-order noise which Adam
# turns into +-lr steps on noise-driven entries, so the yardstick is NCCL against NCCL re-run
trained = {}
for mode in ('peer', 'nccl', 'nccl_again'):
  os.environ['IBC_PEER_ALLREDUCE'] = '1' if mode == 'peer' else '0'
  torch.manual_seed(21)
  net2, agent2, _ = bench.build_agent(dict(bench.CFG, batch=64), torch.device('cuda', rank), seed=9)
  for i in range(8):
    o2, a2 = bench.synthetic_batch(dict(bench.CFG, batch=64), 64, 300 + 10 * i + rank,
                                   device=torch.device('cuda', rank))
    info = agent2.train(((o2, a2), ()))
  agent2.flush_checks()
  used = getattr(agent2, '_peer', None) is not None
  if used != (mode == 'peer'):
    ok = False
    print('PEER PATH NOT TAKEN' if mode == 'peer' else 'PEER PATH TAKEN UNDER IBC_PEER_ALLREDUCE=0', rank, flush=True)
  if used:
    agent2._peer.check()
  trained[mode] = net2.params.detach().clone()
os.environ.pop('IBC_PEER_ALLREDUCE', None)
d = (trained['peer'] - trained['nccl']).abs()
d0 = (trained['nccl_again'] - trained['nccl']).abs()
stats = lambda x: (float(x.median()), float(x.mean()), float(x.max()))
if rank == 0:
  print('PEER_VS_NCCL median/mean/max', stats(d), 'NCCL_VS_NCCL', stats(d0), flush=True)
if not (stats(d)[1] <= 2.0 * stats(d0)[1] + 1e-6 and stats(d)[2] <= 2.0 * stats(d0)[2] + 1e-5):
  ok = False
  print('PEER TRAINING MISMATCH', rank, stats(d), stats(d0), flush=True)
both = [torch.empty_like(trained['peer']) for _ in range(world)]
dist.all_gather(both, trained['peer'])
if not all(torch.equal(both[0], x) for x in both):      # the replicas stay bit-identical
  ok = False
  print('PEER REPLICAS DIVERGED', rank, flush=True)
t = torch.tensor([1.0 if ok else 0.0], device='cuda')
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
  print('DP_PARITY', 'ok' if float(t) == 1.0 else 'failed', flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def _run(script, nproc, tmp_path, timeout=600):
  path = str(tmp_path / 'worker.py')
  with open(path, 'w') as f:
    f.write(script)
  port = 29500 + (os.getpid() % 2000)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc),
         '--master-addr', '127.0.0.1', '--master-port', str(port), path]
  return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_rank_cuda_gradients_match_one_rank(tmp_path):
  out = _run(_WORKER % {'root': ROOT}, 2, tmp_path)
  assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
  assert 'DP_PARITY ok' in out.stdout, out.stdout[-2000:]
