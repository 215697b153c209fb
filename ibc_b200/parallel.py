"""Data-parallel plumbing: one process per GPU, gradients summed
with one all-reduce of the flat fp32 vector (NCCL over NVLink on GPUs, gloo in
the CPU tests).  The loss is already scaled by the GLOBAL batch inside
`ibc_train_grads` (common.aggregate_losses, ibc_agent.py:246-250), so a SUM
reproduces the single-process gradient of the concatenated batch.

Inference shards by OBSERVATION (`sharded_policy_action`): every rank runs the whole
sampler (all N candidates, softmax / CDF local) for its slice of the observation batch; the
only exchange is the optional gather of the A floats per observation.  The reference has
only the degenerate form (run everywhere, keep `local_results[0]`,
ibc/utils/strategy_policy.py:55-68)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def num_replicas() -> int:
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def replica_rank() -> int:
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


_side_streams = {}


def _side_stream(device) -> 'torch.cuda.Stream':
  """One auxiliary stream per device (early loss read-out: ibc_agent._early_loss)."""
  key = torch.device(device).index
  if key not in _side_streams:
    _side_streams[key] = torch.cuda.Stream(device=device)
  return _side_streams[key]


def allreduce_gradients(flat_grads: torch.Tensor) -> torch.Tensor:
  """In-place SUM over replicas of the flat gradient vector."""
  if num_replicas() > 1:
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
  return flat_grads


def per_replica_batch_size(batch_size: int) -> int:
  """train_eval.py:163: per_replica_batch_size = batch_size // num_replicas."""
  return batch_size // num_replicas()


def shard_observations(num_observations: int):
  """Contiguous slice of observations this rank evaluates at inference
  (independent units, no collective)."""
  n, r = num_replicas(), replica_rank()
  per = (num_observations + n - 1) // n
  lo = min(num_observations, r * per)
  return lo, min(num_observations, lo + per)


def sharded_policy_action(policy, time_step, policy_state=(), gather=True, **kwargs):
  """`policy.action` on this rank's contiguous slice of the observation batch.

  Returns (actions, (lo, hi)): with `gather` the [B, A] actions of all observations on every
  rank (one all_gather of A floats per observation), otherwise this rank's [hi - lo, A].
  With one replica this is `policy.action(time_step).action`.  The observation batch is any
  nest of tensors with a common leading dimension (ibc_policy.py:240-262)."""
  from ibc_b200.ibc import specs
  first = specs.flatten(time_step.observation)[0]
  b = int(first.shape[0])
  lo, hi = shard_observations(b)
  n = num_replicas()
  if n == 1:
    return policy.action(time_step, policy_state, **kwargs).action, (0, b)
  local = None
  if hi > lo:
    obs = specs.map_structure(lambda t: t[lo:hi], time_step.observation)
    local_ts = specs.TimeStep(time_step.step_type, time_step.reward, time_step.discount, obs)
    local = policy.action(local_ts, policy_state, **kwargs).action
  if not gather:
    return local, (lo, hi)
  per = (b + n - 1) // n
  a_dim = int(policy.action_spec.shape[-1])
  device = local.device if local is not None else torch.device('cuda', torch.cuda.current_device())
  buf = torch.zeros(per, a_dim, device=device, dtype=torch.float32)
  if local is not None:
    buf[:hi - lo] = local.reshape(hi - lo, a_dim)
  out = torch.empty(n * per, a_dim, device=device, dtype=torch.float32)
  dist.all_gather_into_tensor(out, buf)
  return out[:b], (lo, hi)
