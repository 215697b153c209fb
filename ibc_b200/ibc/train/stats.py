"""Dataset statistics and the normalisation layers built from them
(ibc/train/stats.py:30-310 of the reference), on NumPy / torch.

* `ChanRunningStatistics`: single-pass mean / population variance with Chan's
  parallel update (stats.py:30-83; tf-agents' parallel_variance_calculation).
* `compute_dataset_statistics`: per-key observation statistics over every time
  step of `num_samples` examples, action statistics and per-dimension action
  min / max (stats.py:109-232).
* Layers (stats.py:235-310): Std(De)Normalization with `max(std, eps)`,
  MinMax(De)Normalization to [-1, 1]; callables on torch tensors (any device).
"""
from __future__ import annotations

import numpy as np
import torch

from ibc_b200 import gin_lite as gin

EPS = np.finfo(np.float32).eps


class ChanRunningStatistics:
  """stats.py:30-83."""

  def __init__(self, initial_mean):
    initial_mean = np.asarray(initial_mean)
    self._n = 0
    self._mean = (initial_mean.reshape(-1, initial_mean.shape[-1])[0] if initial_mean.ndim > 1
                  else initial_mean) * 0
    self._m2 = 0.0

  def update_running_statistics(self, sample):
    sample = np.asarray(sample, np.float64)
    if sample.ndim > 1:
      sample = sample.reshape(-1, sample.shape[-1])
      n_b = sample.shape[0]
      avg_b = sample.mean(axis=0)
      m2_b = sample.var(axis=0) * n_b
    else:
      n_b, avg_b, m2_b = 1, sample, 0.0
    n_a, avg_a, m2_a = self._n, self._mean, self._m2
    n_ab = n_a + n_b
    delta = avg_b - avg_a
    self._mean = avg_a + delta * (n_b / n_ab)
    self._m2 = m2_a + m2_b + delta * delta * (n_a * n_b / n_ab)
    self._n = n_ab

  @property
  def mean(self):
    return self._mean

  @property
  def variance(self):
    return self._m2 / self._n

  @property
  def std(self):
    return np.sqrt(self.variance)

  @property
  def n(self):
    return self._n


def _as_like(x, ref):
  return torch.as_tensor(x, dtype=torch.float32, device=ref.device)


class IdentityLayer:
  def __call__(self, x, **kwargs):
    return x.to(torch.float32)


class StdNormalizationLayer:
  """(x - mean) / max(std, eps)   (stats.py:247-258)."""

  def __init__(self, mean, std):
    self._mean = np.asarray(mean, np.float32)
    self._std = np.asarray(std, np.float32)

  def __call__(self, vector, **kwargs):
    vector = vector.to(torch.float32)
    return (vector - _as_like(self._mean, vector)) / _as_like(np.maximum(self._std, EPS), vector)


class StdDenormalizationLayer:
  """stats.py:261-275."""

  def __init__(self, mean, std):
    self._mean = np.asarray(mean, np.float32)
    self._std = np.asarray(std, np.float32)

  def __call__(self, vector, mean_offset=True, **kwargs):
    vector = vector.to(torch.float32)
    result = vector * _as_like(np.maximum(self._std, EPS), vector)
    if mean_offset:
      result = result + _as_like(self._mean, vector)
    return result


class MinMaxLayer:
  """stats.py:278-288."""

  def __init__(self, vmin, vmax):
    self._min = np.asarray(vmin, np.float32)
    self._max = np.asarray(vmax, np.float32)
    self._mean_range = (self._min + self._max) / 2.0
    self._half_range = np.maximum(0.5 * (self._max - self._min), EPS)


class MinMaxNormalizationLayer(MinMaxLayer):
  """Maps to [-1, 1]   (stats.py:291-297)."""

  def __call__(self, vector, **kwargs):
    vector = vector.to(torch.float32)
    return (vector - _as_like(self._mean_range, vector)) / _as_like(self._half_range, vector)


class MinMaxDenormalizationLayer(MinMaxLayer):
  """stats.py:300-306."""

  def __call__(self, vector, **kwargs):
    vector = vector.to(torch.float32)
    return vector * _as_like(self._half_range, vector) + _as_like(self._mean_range, vector)


class NestMap:
  """Applies a dict of layers key by key (ibc/utils/nest_map.py)."""

  def __init__(self, layers: dict):
    self.layers = dict(layers)

  def __call__(self, nest, **kwargs):
    return {k: self.layers[k](v, **kwargs) for k, v in nest.items()}


@gin.configurable
def compute_dataset_statistics(observations, actions, num_samples, nested_obs=True,
                               nested_actions=False, min_max_actions=False,
                               use_sqrt_std=False):
  """stats.py:109-232 over an in-memory dataset.

  observations: dict key -> [S, T, d] (nested_obs) or one [S, T, d] array;
  actions [S, A].  Uses the first `num_samples` examples; every time step of an
  observation window is a sample of its statistics (stats.py:176-179).
  Returns (obs_norm_layer, act_norm_layer, act_denorm_layer, min_actions, max_actions).
  """
  if nested_actions:
    raise NotImplementedError('nested actions are not used by any EBM config')
  if isinstance(observations, dict):
    if not nested_obs and len(observations) > 1:
      raise ValueError('Found too many observations, make sure you set `nested=True` or '
                       'you flatten them.')
    obs_items = [(k, np.asarray(observations[k])) for k in sorted(observations)]
  else:
    obs_items = [(None, np.asarray(observations))]
  actions = np.asarray(actions)
  n = min(int(num_samples), actions.shape[0])
  get = np.sqrt if use_sqrt_std else (lambda x: x)

  obs_layers = {}
  for key, arr in obs_items:
    stat = ChanRunningStatistics(arr[0, 0])
    for i in range(0, n, 1024):          # chunked: Chan's update is exact for any split
      stat.update_running_statistics(arr[i:min(n, i + 1024)])
    assert stat.n > 2
    obs_layers[key] = StdNormalizationLayer(mean=stat.mean, std=get(stat.std))
  act_stat = ChanRunningStatistics(actions[0])
  act_stat.update_running_statistics(actions[:n])
  min_actions = actions[:n].min(axis=0).astype(np.float32)
  max_actions = actions[:n].max(axis=0).astype(np.float32)
  if min_max_actions:
    act_norm = MinMaxNormalizationLayer(vmin=min_actions, vmax=max_actions)
    act_denorm = MinMaxDenormalizationLayer(vmin=min_actions, vmax=max_actions)
  else:
    act_norm = StdNormalizationLayer(mean=act_stat.mean, std=get(act_stat.std))
    act_denorm = StdDenormalizationLayer(mean=act_stat.mean, std=get(act_stat.std))
  if nested_obs and isinstance(observations, dict):
    obs_norm = NestMap(obs_layers)
  else:
    obs_norm = obs_layers[obs_items[0][0]]
  return obs_norm, act_norm, act_denorm, min_actions, max_actions
