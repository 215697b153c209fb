"""ibc/train/get_sampling_spec.py:24-75: bounds for the uniform negative actions = data
min / max widened by `uniform_boundary_buffer` of the range, clipped to the environment's
action bounds, mapped through the action normaliser."""
from __future__ import annotations

import numpy as np
import torch

from ibc_b200.ibc import specs


def get_sampling_spec(action_tensor_spec, min_actions, max_actions, uniform_boundary_buffer,
                      act_norm_layer):
  min_action = np.asarray(min_actions, np.float32).copy()
  max_action = np.asarray(max_actions, np.float32).copy()
  action_range = max_action - min_action
  min_action -= action_range * uniform_boundary_buffer
  max_action += action_range * uniform_boundary_buffer
  min_action = np.maximum(action_tensor_spec.minimum, min_action)
  max_action = np.minimum(action_tensor_spec.maximum, max_action)
  lo = act_norm_layer(torch.as_tensor(min_action)[None])[0].numpy()
  hi = act_norm_layer(torch.as_tensor(max_action)[None])[0].numpy()
  return specs.BoundedTensorSpec(action_tensor_spec.shape, action_tensor_spec.dtype,
                                 minimum=lo, maximum=hi, name=action_tensor_spec.name)
