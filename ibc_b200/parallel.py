"""Data-parallel plumbing: one process per GPU, gradients summed
with one all-reduce of the flat fp32 vector (NCCL over NVLink on GPUs, gloo in
the CPU tests).  The loss is already scaled by the GLOBAL batch inside
`ibc_train_grads` (common.aggregate_losses, ibc_agent.py:246-250), so a SUM
reproduces the single-process gradient of the concatenated batch."""
from __future__ import annotations

import torch
import torch.distributed as dist


def num_replicas() -> int:
  return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def replica_rank() -> int:
  return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


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
