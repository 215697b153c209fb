"""ibc/train/get_normalizers.py:51-110: statistics -> normaliser layers + the function
that normalises a training batch."""
from __future__ import annotations

import collections

from ibc_b200 import gin_lite as gin
from ibc_b200.ibc.train import stats

NormalizationInfo = collections.namedtuple(
    'NormalizationInfo',
    ('obs_norm_layer', 'act_norm_layer', 'act_denorm_layer', 'min_actions', 'max_actions'))


@gin.configurable
def get_normalizers(observations, actions, batch_size, env_name=None, nested_obs=False,
                    nested_actions=False, num_batches=100, num_samples=None):
  if num_samples is None:
    num_samples = num_batches * batch_size          # get_normalizers.py:72-73
  (obs_norm_layer, act_norm_layer, act_denorm_layer, min_actions, max_actions) = (
      stats.compute_dataset_statistics(observations, actions, num_samples=num_samples,
                                       nested_obs=nested_obs, nested_actions=nested_actions))

  def norm_train_data_fn(obs_and_act, nothing=()):
    obs, act = obs_and_act
    return (obs_norm_layer(obs), act_norm_layer(act)), nothing

  info = NormalizationInfo(obs_norm_layer, act_norm_layer, act_denorm_layer, min_actions,
                           max_actions)
  return info, norm_train_data_fn
