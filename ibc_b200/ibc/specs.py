"""Minimal stand-ins for the tf-agents types the reference's agent/policy
surface uses (tf-agents is not installed): tensor specs, TimeStep, PolicyStep,
LossInfo and nest_utils.tile_batch."""
from __future__ import annotations

import collections
from typing import Sequence

import numpy as np
import torch


class TensorSpec:
  def __init__(self, shape: Sequence[int], dtype=torch.float32, name=None):
    self.shape = tuple(shape)
    self.dtype = dtype
    self.name = name

  def __repr__(self):
    return 'TensorSpec(shape=%s, dtype=%s, name=%r)' % (self.shape, self.dtype, self.name)


class BoundedTensorSpec(TensorSpec):
  def __init__(self, shape, dtype=torch.float32, minimum=-1.0, maximum=1.0, name=None):
    super().__init__(shape, dtype, name)
    self.minimum = np.broadcast_to(np.asarray(minimum, dtype=np.float32), self.shape).copy()
    self.maximum = np.broadcast_to(np.asarray(maximum, dtype=np.float32), self.shape).copy()


TimeStep = collections.namedtuple('TimeStep', ['step_type', 'reward', 'discount', 'observation'])
PolicyStep = collections.namedtuple('PolicyStep', ['action', 'state', 'info'])
PolicyStep.__new__.__defaults__ = ((), ())
LossInfo = collections.namedtuple('LossInfo', ['loss', 'extra'])


def restart(observation):
  """ts.restart-like helper for tests/benchmarks."""
  return TimeStep(step_type=0, reward=0.0, discount=1.0, observation=observation)


def map_structure(fn, nest):
  if isinstance(nest, dict):
    return {k: map_structure(fn, nest[k]) for k in nest}
  if isinstance(nest, (list, tuple)) and not hasattr(nest, '_fields'):
    return type(nest)(map_structure(fn, x) for x in nest)
  return fn(nest)


def flatten(nest):
  """tf.nest.flatten: dict values in SORTED key order (mlp_ebm.py:78)."""
  if isinstance(nest, dict):
    out = []
    for k in sorted(nest):
      out.extend(flatten(nest[k]))
    return out
  if isinstance(nest, (list, tuple)):
    out = []
    for x in nest:
      out.extend(flatten(x))
    return out
  return [nest]


def tile_batch(nest, multiplier: int):
  """nest_utils.tile_batch: every batch entry repeated `multiplier` times
  contiguously (row r <-> entry r // multiplier)."""
  return map_structure(lambda t: torch.repeat_interleave(t, multiplier, dim=0), nest)


def flatten_observation(obs) -> torch.Tensor:
  """[R, T, d_k] per key -> [R, T * sum d_k] (mlp_ebm.py:78-82)."""
  parts = flatten(obs)
  x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=-1)
  return x.reshape(x.shape[0], -1)
