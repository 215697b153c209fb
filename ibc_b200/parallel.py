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


class _RawDeviceMemory:
  """__cuda_array_interface__ over memory the library allocated (kept alive by `owner`)."""

  def __init__(self, ptr, numel, typestr, owner):
    self.owner = owner
    self.__cuda_array_interface__ = {'shape': (numel,), 'typestr': typestr, 'data': (ptr, False),
                                     'version': 2, 'strides': None}


def _device_view(ptr, numel, dtype, device, owner) -> torch.Tensor:
  typestr = {torch.float32: '<f4', torch.int32: '<i4'}[dtype]
  return torch.as_tensor(_RawDeviceMemory(ptr, numel, typestr, owner), device=device)


class PeerReducer:
  """Gradient all-reduce + Adam in one launch over NVLink peer memory (ibc_peer_allreduce_adam).

  Every rank owns two gradient buffers (step parity) and a flag array; their CUDA IPC handles are
  exchanged once through the process group, so that each rank's kernel can read every peer's
  gradient directly.  One node, 2..8 ranks, one GPU per rank.  `create` returns None when the
  set-up is not possible (single replica, gloo / CPU, peer access missing, IBC_PEER_ALLREDUCE=0):
  the caller then keeps the NCCL all-reduce."""

  def __init__(self, numel: int, device):
    import ctypes
    from ibc_b200 import _lib
    self._lib = _lib
    lib = _lib.load()
    self.rank, self.world = replica_rank(), num_replicas()
    self.numel = numel
    # one allocation per rank, made by the library (its IPC handle names it at offset 0):
    # [gradient, even steps | gradient, odd steps | flags (64 ints)]
    stride = (numel * 4 + 255) // 256 * 256
    self._bytes = 2 * stride + 256
    # Every rank runs the SAME collectives whatever fails locally (one all_gather_object here, one
    # all_reduce in `create`): a rank that cannot allocate or map still takes part, publishes
    # None and raises afterwards, so that nobody is left waiting in a collective.
    self._base, self._opened = None, []
    with torch.cuda.device(device):
      base = ctypes.c_void_p()
      handle = ctypes.create_string_buffer(64)
      mine, failure = None, None
      try:
        _lib.check(lib.ibc_peer_alloc(self._bytes, ctypes.byref(base), handle))
        self._base = base.value
        mine = bytes(handle.raw)
      except Exception as e:                    # noqa: BLE001
        failure = e
      everyone = [None] * self.world
      dist.all_gather_object(everyone, mine)
      if failure is not None or any(hd is None for hd in everyone):
        self._release()
        raise failure or RuntimeError('a peer could not allocate its buffers')
      bases = []
      try:
        for r, hd in enumerate(everyone):
          if r == self.rank:
            bases.append(self._base)
            continue
          q = ctypes.c_void_p()
          _lib.check(lib.ibc_peer_open(ctypes.create_string_buffer(hd, 64), ctypes.byref(q)))
          self._opened.append(q.value)
          bases.append(q.value)
      except Exception:
        self._release()
        raise
    self.bufs = [_device_view(self._base + k * stride, numel, torch.float32, device, self) for k in range(2)]
    arr = ctypes.c_void_p * 8
    pad = [None] * (8 - self.world)
    self._g = [arr(*([x + k * stride for x in bases] + pad)) for k in range(2)]
    self._f = arr(*([x + 2 * stride for x in bases] + pad))
    # the reduced last gradient entry lands in pinned HOST memory (device-addressable): the
    # caller's finiteness check reads it without a copy or a probe kernel
    self.tail = torch.zeros(1, dtype=torch.float32, pin_memory=True)
    self.status = torch.zeros(1, device=device, dtype=torch.int32)
    self.epoch = 0
    torch.cuda.synchronize(device)
    # (no barrier here: `create` ends with an all_reduce on every rank, which is one)

  def _release(self):
    lib = self._lib.load()
    for q in self._opened:
      lib.ibc_peer_close(q, 0)
    if self._base:
      lib.ibc_peer_close(self._base, 1)
    self._base, self._opened, self.bufs = None, [], []

  def close(self):
    """Unmaps the peers' buffers and frees this rank's (all ranks, after their last step)."""
    if getattr(self, '_base', None):
      lib = self._lib.load()
      torch.cuda.synchronize()
      dist.barrier()                 # no peer still reads this rank's buffers
      self._release()

  @classmethod
  def create(cls, numel: int, device):
    import os
    if num_replicas() < 2 or num_replicas() > 8 or os.environ.get('IBC_PEER_ALLREDUCE', '1') == '0':
      return None
    if torch.device(device).type != 'cuda' or dist.get_backend() != 'nccl':
      return None
    ok = torch.ones(1, device=device)
    try:
      red = cls(numel, device)
    except Exception as e:                      # noqa: BLE001 -- any failure: keep NCCL
      red = None
      ok.zero_()
      import warnings
      warnings.warn('peer-memory all-reduce unavailable (%s): using NCCL' % (e,))
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # all ranks take the same path (and a barrier)
    if float(ok) > 0:
      return red
    if red is not None:                          # a peer failed: give the buffers back
      red._release()
    return None

  def buffer(self, step: int) -> torch.Tensor:
    """The gradient buffer of this step (its parity): what train_grads writes."""
    return self.bufs[step & 1]

  def reduce_and_adam(self, step: int, params, m, v, lr_t, beta1, beta2, epsilon):
    """Sum of all replicas' bufs[step & 1], Adam on this replica's (params, m, v); self.tail
    holds the reduced last gradient entry afterwards."""
    import ctypes
    self.epoch += 1
    lib = self._lib.load()
    self._lib.check(lib.ibc_peer_allreduce_adam(
        self._g[step & 1], self._f, self.rank, self.world, self.epoch, self.numel,
        ctypes.c_void_p(params.data_ptr()), ctypes.c_void_p(m.data_ptr()), ctypes.c_void_p(v.data_ptr()),
        float(lr_t), float(beta1), float(beta2), float(epsilon), ctypes.c_void_p(self.tail.data_ptr()),
        ctypes.c_void_p(self.status.data_ptr()),
        ctypes.c_void_p(torch.cuda.current_stream(params.device).cuda_stream)))

  def check(self):
    """Raises if a reduction timed out waiting for a peer (device sync)."""
    st = int(self.status.item())
    if st:
      raise RuntimeError('peer-memory all-reduce failed: kernel status %d (a peer did not arrive)' % st)


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
