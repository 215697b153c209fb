"""Optimizer + agent factory of the reference `ibc/train/get_agent.py:49-90`:
Keras-2.6 Adam (beta 0.9/0.999, epsilon 1e-7, lr_t = lr * sqrt(1-b2^t)/(1-b1^t),
theta -= lr_t * m / (sqrt(v) + eps)) with ExponentialDecay(lr, decay_steps,
0.99) evaluated continuously.  The update itself is `ibc_adam_step`."""
from __future__ import annotations

import math

import torch

from ibc_b200 import engine as eng
from ibc_b200.ibc.agents import ibc_agent


class ExponentialDecay:
  def __init__(self, initial_learning_rate, decay_steps, decay_rate, staircase=False):
    self.initial_learning_rate = initial_learning_rate
    self.decay_steps = decay_steps
    self.decay_rate = decay_rate
    self.staircase = staircase

  def __call__(self, step):
    p = step / self.decay_steps
    if self.staircase:
      p = math.floor(p)
    return self.initial_learning_rate * self.decay_rate ** p


class Adam:
  """tf.keras.optimizers.Adam look-alike over one flat parameter tensor."""

  def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    self.learning_rate = learning_rate
    self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
    self.iterations = 0
    self.m = None
    self.v = None

  def _lr(self):
    lr = self.learning_rate
    return float(lr(self.iterations)) if callable(lr) else float(lr)

  def lr_t(self):
    t = self.iterations + 1
    return self._lr() * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

  def apply_gradients(self, network, grads: torch.Tensor):
    params = network.params
    if self.m is None:
      self.m = torch.zeros_like(params)
      self.v = torch.zeros_like(params)
    eng.adam_step(params, grads, self.m, self.v, self.lr_t(), self.beta_1, self.beta_2,
                  self.epsilon)
    self.iterations += 1
    network.params_changed()

  def apply_gradients_peer(self, network, reducer, step: int):
    """The same update with the data-parallel gradient sum inside the launch
    (parallel.PeerReducer: every replica reads its peers' gradients over NVLink)."""
    params = network.params
    if self.m is None:
      self.m = torch.zeros_like(params)
      self.v = torch.zeros_like(params)
    reducer.reduce_and_adam(step, params, self.m, self.v, self.lr_t(), self.beta_1, self.beta_2,
                            self.epsilon)
    self.iterations += 1
    network.params_changed()


def get_agent(loss_type, time_step_tensor_spec, action_tensor_spec, action_sampling_spec,
              obs_norm_layer, act_norm_layer, act_denorm_layer, learning_rate, use_warmup,
              cloning_network, train_step, decay_steps):
  """get_agent.py:49-90 (EBM branch)."""
  if use_warmup:
    raise NotImplementedError('WarmupSchedule is only used by the MSE/MDN transformers')
  if loss_type != 'ebm':
    raise ValueError("Unsupported loss type, can't retrieve an agent.")
  schedule = ExponentialDecay(learning_rate, decay_steps=decay_steps, decay_rate=0.99)
  optimizer = Adam(learning_rate=schedule)
  return ibc_agent.ImplicitBCAgent(
      time_step_spec=time_step_tensor_spec, action_spec=action_tensor_spec,
      action_sampling_spec=action_sampling_spec, obs_norm_layer=obs_norm_layer,
      act_norm_layer=act_norm_layer, act_denorm_layer=act_denorm_layer,
      cloning_network=cloning_network, optimizer=optimizer, train_step_counter=train_step)
