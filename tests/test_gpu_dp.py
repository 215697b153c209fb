"""Data-parallel parity on the GPU: two NCCL ranks run the CUDA
`ibc_train_grads` on the two halves of a batch with loss scaling by the GLOBAL batch
(common.aggregate_losses, ibc/agents/ibc_agent.py:246-250), all-reduce the flat gradient
(ibc_b200.parallel.allreduce_gradients, the MirroredStrategy sum of
ibc/agents/base_agent.py:50-58), and must reproduce the one-rank gradient of the whole batch:
1e-5 relative in fp32 (reduction order), 2e-3 per tensor in tf32 (the fp16 dY scale of the
block-major backward is taken per rank).  Skipped with fewer than two GPUs; also covers the
observation-sharded policy (row e2): two ranks' actions == the one-rank batched actions."""
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
