"""Samplers of the reference `ibc/agents/mcmc.py`, re-hosted on the B200 engine.

Same names, argument order, defaults and return values as the reference
(file:line cited per function).  Two execution paths, both CUDA-only:
  * `network` is an `ibc_b200` MLPEBM: the whole sampler runs inside the
    library (`ibc_dfo` / `ibc_langevin`, include/ibc_b200.h);
  * any other callable `network((obs, act), training=..., ...) -> (E, state)`
    on CUDA torch tensors (e.g. the mock net of mcmc_test.py:67-81): energies
    and dE/dact come from the callable (torch autograd), every sampler step
    (resample, gather/perturb, Langevin update, softmax) from the library.
Extra keyword-only arguments `uniform_draws=` / `normal_draws=` / `seed=` make
the random draws explicit for parity tests.
"""
from __future__ import annotations

import collections
from typing import Optional

import numpy as np
import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin
from ibc_b200.ibc import specs
from ibc_b200.networks.mlp_ebm import MLPEBM
from ibc_b200.networks.pixel_ebm import PixelEBM

my_range = range


def _next_seed() -> int:
  """Fresh 62-bit seed from torch's default CPU generator (so
  torch.manual_seed makes sampler draws reproducible)."""
  return int(torch.randint(0, 2 ** 62, (1,)).item())


def _as_f32(t, device):
  if not torch.is_tensor(t):
    t = torch.as_tensor(np.asarray(t, dtype=np.float32))
  return t.to(device=device, dtype=torch.float32).contiguous()


def _untile(obs_flat: torch.Tensor, batch_size: int, n: int) -> torch.Tensor:
  """[B*n, O] produced by tile_batch -> [B, O]."""
  if obs_flat.shape[0] == batch_size:
    return obs_flat.contiguous()
  return obs_flat.view(batch_size, n, -1)[:, 0].contiguous()


def categorical_bincount(count, log_ps, n, *, uniform_draws=None, seed=None):
  """mcmc.py:32-55: counts [B, n] of `count` categorical draws per row."""
  log_ps = log_ps.to(torch.float32).contiguous()
  assert log_ps.shape[1] == n
  b = log_ps.shape[0]
  if uniform_draws is None:
    g = torch.Generator(device=log_ps.device)
    g.manual_seed(_next_seed() if seed is None else int(seed))
    uniform_draws = torch.rand(b, count, generator=g, device=log_ps.device, dtype=torch.float64)
  _, _, counts = eng.resample(log_ps, uniform_draws.contiguous(), want_counts=True)
  return counts


def iterative_dfo(network,
                  batch_size,  # B
                  observations,  # B*n x obs_spec or B x obs_spec if late_fusion
                  action_samples,  # B*n x act_spec
                  policy_state,
                  num_action_samples,  # n
                  min_actions,
                  max_actions,
                  temperature=1.0,
                  num_iterations=3,
                  iteration_std=0.33,
                  training=False,
                  late_fusion=False,
                  tfa_step_type=(),
                  *,
                  uniform_draws=None,
                  normal_draws=None,
                  seed=None):
  """mcmc.py:58-158.  Returns (probs [B*n], action_samples [B*n, A], state).

  uniform_draws [num_iterations, B, n] float64, normal_draws
  [num_iterations-1, B*n, A] float32."""
  n = int(num_action_samples)
  batch_size = int(batch_size)
  device = action_samples.device
  min_actions = _as_f32(min_actions, device)
  max_actions = _as_f32(max_actions, device)
  seed = _next_seed() if seed is None else int(seed)
  if isinstance(network, MLPEBM) and not late_fusion:
    obs = _untile(network._flat_obs(observations), batch_size, n)
    actions = action_samples.to(torch.float32).contiguous().clone()
    probs = network.engine.dfo(obs, actions, n, min_actions, max_actions, temperature,
                               num_iterations, iteration_std, uniform_draws, normal_draws, seed)
    return probs, actions, policy_state

  if isinstance(network, PixelEBM) and late_fusion:
    # embed observations once (:98-102), then the whole loop runs in the library
    enc = network.encode(observations, training=training)
    actions = action_samples.to(torch.float32).contiguous().clone()
    probs = network.engine.dfo(enc, actions, n, min_actions, max_actions, temperature,
                               num_iterations, iteration_std, uniform_draws, normal_draws, seed)
    return probs, actions, policy_state

  # generic energy callable -------------------------------------------------
  obs_encodings = None
  if late_fusion:
    obs_encodings = network.encode(observations, training=training)        # :98-102
    obs_encodings = specs.tile_batch(obs_encodings, n)
  gen = torch.Generator(device=device)
  gen.manual_seed(seed)

  def update_selected_actions(samples, policy_state, it):
    with torch.no_grad():
      if late_fusion:
        net_logits, new_state = network((observations, samples), step_type=tfa_step_type,
                                        training=training, network_state=policy_state,
                                        observation_encoding=obs_encodings)
      else:
        net_logits, new_state = network((observations, samples), step_type=tfa_step_type,
                                        network_state=policy_state, training=training)
    log_probs = (net_logits.reshape(batch_size, n) / temperature).to(torch.float32).contiguous()
    if uniform_draws is not None:
      u = uniform_draws[it].contiguous()
    else:
      u = torch.rand(batch_size, n, generator=gen, device=device, dtype=torch.float64)
    rows, _, _ = eng.resample(log_probs, u)
    return log_probs, rows, new_state

  samples = action_samples.to(torch.float32).contiguous()
  log_probs, rows, new_state = update_selected_actions(samples, policy_state, 0)
  std = float(iteration_std)
  for it in my_range(1, num_iterations):
    nd = None if normal_draws is None else normal_draws[it - 1].contiguous()
    samples = eng.gather_perturb(samples, rows, std, min_actions, max_actions, nd, seed + it)
    log_probs, rows, new_state = update_selected_actions(samples, new_state, it)
    std *= 0.5                                                               # :153
  samples = eng.gather_perturb(samples, rows, 0.0, min_actions, max_actions)
  probs = eng.softmax_rows(log_probs, 1.0).reshape(-1)                       # :155-156
  return probs, samples, new_state


def gradient_wrt_act(energy_network,
                     observations,
                     actions,
                     training,
                     network_state,
                     tfa_step_type,
                     apply_exp,
                     obs_encoding=None):
  """mcmc.py:161-191: (-dE/dact, energies)."""
  if isinstance(energy_network, MLPEBM) and obs_encoding is None:
    if apply_exp:
      raise NotImplementedError('apply_exp is only available through langevin_actions_given_obs')
    obs = energy_network._flat_obs(observations)
    act = actions.to(torch.float32).contiguous()
    e, g = energy_network.engine.energy_dact(obs, act, 1)
    return -g, e
  if isinstance(energy_network, PixelEBM) and not apply_exp:
    # late fusion: the value head on a prior encoding (pixel_ebm.py:113-125); without one
    # the network encodes and tiles itself (pixel_ebm.py:107-112, gradient_loss.py:38-45)
    act = actions.to(energy_network.device, torch.float32).contiguous()
    if obs_encoding is None:
      enc = energy_network.encode(observations, training)
      e, g = energy_network.engine.energy_dact(enc, act, act.shape[0] // enc.shape[0])
    else:
      enc = obs_encoding.to(energy_network.device, torch.float32).contiguous()
      if enc.shape[0] != act.shape[0]:
        raise ValueError('observation_encoding must be tiled to the actions')
      e, g = energy_network.engine.energy_dact(enc, act, 1)
    return -g, e
  a = actions.detach().clone().requires_grad_(True)
  with torch.enable_grad():
    if obs_encoding is not None:
      energies, _ = energy_network((observations, a), training=training,
                                   network_state=network_state, step_type=tfa_step_type,
                                   observation_encoding=obs_encoding)
    else:
      energies, _ = energy_network((observations, a), training=training,
                                   network_state=network_state, step_type=tfa_step_type)
    if apply_exp:
      energies = torch.exp(energies)
    (g,) = torch.autograd.grad(energies.sum(), a)
  return g * -1.0, energies.detach()


def compute_grad_norm(grad_norm_type, de_dact):
  """mcmc.py:194-206."""
  if grad_norm_type is not None:
    grad_norm_type_to_ord = {'1': 1, '2': 2, 'inf': float('inf')}
    grad_type = grad_norm_type_to_ord[grad_norm_type]
    return torch.linalg.norm(de_dact, ord=grad_type, dim=1)
  return torch.zeros_like(de_dact[:, 0])


def langevin_step(energy_network,
                  observations,
                  actions,
                  training,
                  policy_state,
                  tfa_step_type,
                  noise_scale,
                  grad_clip,
                  delta_action_clip,
                  stepsize,
                  apply_exp,
                  min_actions,
                  max_actions,
                  grad_norm_type,
                  obs_encoding,
                  *,
                  normal_draws=None,
                  seed=None):
  """mcmc.py:209-264.  Returns (actions, energies, grad_norms)."""
  de_dact, energies = gradient_wrt_act(energy_network, observations, actions, training,
                                       policy_state, tfa_step_type, apply_exp, obs_encoding)
  device = actions.device
  new_actions = actions.to(torch.float32).contiguous().clone()
  # the update kernel takes +dE/dact
  norms = eng.langevin_update(new_actions, (-de_dact).contiguous(), float(stepsize),
                              _as_f32(min_actions, device), _as_f32(max_actions, device),
                              noise_scale, grad_clip, delta_action_clip,
                              grad_norm_type, normal_draws,
                              _next_seed() if seed is None else int(seed))
  return new_actions, energies, norms


class ExponentialSchedule:
  """Exponential learning rate schedule for Langevin sampler (mcmc.py:267-278)."""

  def __init__(self, init, decay):
    self._decay = decay
    self._latest_lr = init

  def get_rate(self, index):
    """Get learning rate. Assumes calling sequentially."""
    del index
    self._latest_lr *= self._decay
    return self._latest_lr


class PolynomialSchedule:
  """Polynomial learning rate schedule for Langevin sampler (mcmc.py:281-294)."""

  def __init__(self, init, final, power, num_steps):
    self._init = init
    self._final = final
    self._power = power
    self._num_steps = num_steps

  def get_rate(self, index):
    """Get learning rate for index."""
    return ((self._init - self._final) *
            ((1 - (float(index) / float(self._num_steps - 1))) ** (self._power))
            ) + self._final


ChainData = collections.namedtuple('ChainData', ['actions', 'energies', 'grad_norms'])


def update_chain_data(num_iterations, step_index, actions, energies, grad_norms,
                      full_chain_actions, full_chain_energies, full_chain_grad_norms):
  """mcmc.py:297-327 (one-hot accumulate == indexed write)."""
  full_chain_actions[step_index] = actions
  full_chain_energies[step_index] = energies
  full_chain_grad_norms[step_index] = grad_norms
  return full_chain_actions, full_chain_energies, full_chain_grad_norms


@gin.configurable
def langevin_actions_given_obs(
    energy_network,
    observations,  # B*n x obs_spec or B x obs_spec if late_fusion
    action_samples,  # B*n x act_spec
    policy_state,
    min_actions,
    max_actions,
    num_action_samples,
    num_iterations=25,
    training=False,
    tfa_step_type=(),
    sampler_stepsize_init=1e-1,
    sampler_stepsize_decay=0.8,  # if using exponential langevin rate.
    noise_scale=1.0,
    grad_clip=None,
    delta_action_clip=0.1,
    stop_chain_grad=True,
    apply_exp=False,
    use_polynomial_rate=True,  # default is exponential
    sampler_stepsize_final=1e-5,  # if using polynomial langevin rate.
    sampler_stepsize_power=2.0,  # if using polynomial langevin rate.
    return_chain=False,
    grad_norm_type='inf',
    late_fusion=False,
    *,
    normal_draws=None,
    seed=None,
    observation_encoding=None):
  """mcmc.py:330-425.  normal_draws [num_iterations, B*n, A] float32.
  observation_encoding (late fusion on a PixelEBM only): the [B, 256] encoding of
  `observations` when the caller already has it (skips the encode of :384-388)."""
  if not stop_chain_grad:
    raise NotImplementedError('stop_chain_grad=False (gradients through the chain) is not built; '
                              'no shipped gin uses it')
  device = action_samples.device
  min_actions = _as_f32(min_actions, device)
  max_actions = _as_f32(max_actions, device)
  n = int(num_action_samples)
  seed = _next_seed() if seed is None else int(seed)
  actions = action_samples.to(torch.float32).contiguous().clone()
  r = actions.shape[0]

  if isinstance(energy_network, MLPEBM) and not late_fusion:
    cfg = eng.langevin_cfg(num_iterations, sampler_stepsize_init, sampler_stepsize_decay,
                           noise_scale, grad_clip, delta_action_clip, use_polynomial_rate,
                           sampler_stepsize_final, sampler_stepsize_power, grad_norm_type,
                           apply_exp)
    obs = _untile(energy_network._flat_obs(observations), r // n, n)
    chain = energy_network.engine.langevin(obs, actions, n, min_actions, max_actions, cfg,
                                           normal_draws, seed, return_chain)
    if return_chain:
      return actions, ChainData(*chain)
    return actions

  if isinstance(energy_network, PixelEBM) and late_fusion and not apply_exp:
    # encode once, then the whole chain on the value head (mcmc.py:384-388)
    cfg = eng.langevin_cfg(num_iterations, sampler_stepsize_init, sampler_stepsize_decay,
                           noise_scale, grad_clip, delta_action_clip, use_polynomial_rate,
                           sampler_stepsize_final, sampler_stepsize_power, grad_norm_type,
                           apply_exp)
    enc = observation_encoding if observation_encoding is not None else \
        energy_network.encode(observations, training=training)
    if enc.shape[0] * n != r:
      raise ValueError('late fusion expects un-tiled observations: %d x %d != %d'
                       % (enc.shape[0], n, r))
    chain = energy_network.engine.langevin(enc, actions, n, min_actions, max_actions, cfg,
                                           normal_draws, seed, return_chain)
    if return_chain:
      return actions, ChainData(*chain)
    return actions

  # generic energy callable -------------------------------------------------
  stepsize = sampler_stepsize_init
  if use_polynomial_rate:
    schedule = PolynomialSchedule(sampler_stepsize_init, sampler_stepsize_final,
                                  sampler_stepsize_power, num_iterations)
  else:
    schedule = ExponentialSchedule(sampler_stepsize_init, sampler_stepsize_decay)
  a_dim = actions.shape[1]
  full_chain_actions = torch.zeros(num_iterations, r, a_dim, device=device)
  full_chain_energies = torch.zeros(num_iterations, r, device=device)
  full_chain_grad_norms = torch.zeros(num_iterations, r, device=device)
  obs_encoding = None
  if late_fusion:
    obs_encoding = energy_network.encode(observations, training=training)    # :384-388
    obs_encoding = specs.tile_batch(obs_encoding, n)
  for step_index in my_range(num_iterations):
    nd = None if normal_draws is None else normal_draws[step_index].contiguous()
    actions, energies, grad_norms = langevin_step(
        energy_network, observations, actions, training, policy_state, tfa_step_type,
        noise_scale, grad_clip, delta_action_clip, stepsize, apply_exp, min_actions,
        max_actions, grad_norm_type, obs_encoding, normal_draws=nd, seed=seed + step_index)
    actions = actions.detach()
    stepsize = schedule.get_rate(step_index + 1)                             # :408
    if return_chain:
      update_chain_data(num_iterations, step_index, actions, energies, grad_norms,
                        full_chain_actions, full_chain_energies, full_chain_grad_norms)
  if return_chain:
    return actions, ChainData(full_chain_actions, full_chain_energies, full_chain_grad_norms)
  return actions


def get_probabilities(energy_network,
                      batch_size,
                      num_action_samples,
                      observations,
                      actions,
                      training,
                      temperature=1.0):
  """mcmc.py:428-441: softmax(E / T) over the n samples, flattened."""
  if isinstance(energy_network, MLPEBM):
    obs = _untile(energy_network._flat_obs(observations), int(batch_size), int(num_action_samples))
    net_logits = energy_network.energy_untiled(obs, actions.to(torch.float32).contiguous(),
                                               int(num_action_samples))
  else:
    with torch.no_grad():
      net_logits, _ = energy_network((observations, actions), training=training)
  net_logits = net_logits.reshape(int(batch_size), int(num_action_samples)).contiguous()
  probs = eng.softmax_rows(net_logits.to(torch.float32), float(temperature))
  return probs.reshape(-1)
