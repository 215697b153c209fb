"""IbcPolicy on the B200 engine -- the reference's
`ibc/agents/ibc_policy.py:132-328` (+ `MappedCategorical` :68-129 and the
tf-agents GreedyPolicy that wraps it).

`policy.action(time_step)` samples N actions uniformly in the sampling bounds,
refines them with DFO and/or Langevin inside libibc_b200.so and returns the
arg-max-probability action.  The reference builds ONE categorical over the
flat [B*N] probabilities, so its greedy action is only meaningful at B=1
(SURVEY.md 8a); `GreedyPolicy` here takes the arg max per observation, which
equals the reference at B=1 (use `MappedCategorical.mode()` for the literal
flat behaviour).
"""
from __future__ import annotations

import numpy as np
import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin
from ibc_b200.ibc import specs
from ibc_b200.ibc.agents import mcmc


class MappedCategorical:
  """Categorical distribution that maps classes to specific values
  (ibc_policy.py:68-129)."""

  def __init__(self, logits=None, probs=None, mapped_values=None):
    if (logits is None) == (probs is None):
      raise ValueError('Must pass probs or logits, but not both.')
    self._probs = probs if probs is not None else torch.softmax(logits, dim=-1)
    self._mapped_values = mapped_values

  @property
  def probs(self):
    return self._probs

  @property
  def mapped_values(self):
    return self._mapped_values

  def mode(self):
    """ONE arg max over the flat probs -> [1, A] (ibc_policy.py:114-117)."""
    mode = torch.argmax(self._probs.reshape(-1))
    return self._mapped_values[mode][None]

  def sample(self, sample_shape=(), seed=None):
    g = None
    if seed is not None:
      g = torch.Generator(device=self._probs.device)
      g.manual_seed(int(seed))
    num = int(np.prod(sample_shape)) if sample_shape else 1
    idx = torch.multinomial(self._probs.reshape(-1), num, replacement=True, generator=g)
    out = self._mapped_values[idx]
    return out.reshape(tuple(sample_shape) + (1,) + tuple(out.shape[1:])) if sample_shape \
        else out


@gin.configurable
class IbcPolicy:
  """Class to build Actor Policies (ibc_policy.py:132-233)."""

  def __init__(self,
               time_step_spec,
               action_spec,
               action_sampling_spec,
               actor_network,
               policy_state_spec=(),
               info_spec=(),
               num_action_samples=2**14,
               clip=True,
               training=False,
               name=None,
               use_dfo=False,
               use_langevin=True,
               inference_langevin_noise_scale=1.0,
               optimize_again=False,
               again_stepsize_init=1e-1,
               again_stepsize_final=1e-5,
               late_fusion=False,
               obs_norm_layer=None,
               act_denorm_layer=None):
    self.time_step_spec = time_step_spec
    self.action_spec = action_spec
    self.policy_state_spec = policy_state_spec
    self.info_spec = info_spec
    self.name = name
    self._action_sampling_spec = action_sampling_spec
    self._num_action_samples = int(num_action_samples)
    self._clip = clip
    self._use_dfo = use_dfo
    self._use_langevin = use_langevin
    self._inference_langevin_noise_scale = inference_langevin_noise_scale
    self._optimize_again = optimize_again
    self._again_stepsize_init = again_stepsize_init
    self._again_stepsize_final = again_stepsize_final
    self._late_fusion = late_fusion
    self._obs_norm_layer = obs_norm_layer
    self._act_denorm_layer = act_denorm_layer
    self._actor_network = actor_network
    self._training = training
    device = getattr(actor_network, 'device', None)
    if device is None:
      device = torch.device('cuda', torch.cuda.current_device())
    self._device = device
    # own copies of the sampling bounds (ibc_policy.py:202-209)
    self._action_sampling_minimum = torch.as_tensor(
        np.asarray(action_sampling_spec.minimum, np.float32)).to(device).reshape(-1)
    self._action_sampling_maximum = torch.as_tensor(
        np.asarray(action_sampling_spec.maximum, np.float32)).to(device).reshape(-1)

  def variables(self):
    return list(getattr(self._actor_network, 'variables', [])) + [
        self._action_sampling_minimum, self._action_sampling_maximum]

  def get_initial_state(self, batch_size=None):
    del batch_size
    return ()

  def _distribution(self, time_step, policy_state, *, draws=None):
    """ibc_policy.py:240-328.  `draws` (optional dict: uniform_u, dfo_uniforms,
    dfo_normals, langevin_normals, again_normals) makes the sampling explicit."""
    draws = draws or {}
    observations = time_step.observation
    if isinstance(observations, dict) and 'rgb' in observations:
      rgb = observations['rgb']
      if rgb.dtype == torch.uint8:                    # convert_image_dtype (:245-247)
        observations = dict(observations)
        observations['rgb'] = rgb.to(torch.float32) / 255.0
    observations = specs.map_structure(
        lambda t: torch.as_tensor(t).to(self._device, non_blocking=True), observations)
    if self._obs_norm_layer is not None:
      observations = self._obs_norm_layer(observations)
    single_obs = specs.flatten(observations)[0]
    batch_size = single_obs.shape[0]
    n = self._num_action_samples
    if self._late_fusion:
      maybe_tiled_obs = observations
    else:
      maybe_tiled_obs = specs.tile_batch(observations, n)
    mn, mx = self._action_sampling_minimum, self._action_sampling_maximum
    u = draws.get('uniform_u')
    action_samples = eng.uniform_actions(batch_size * n, mn, mx, u, mcmc._next_seed())
    probs = 0
    if self._use_dfo:
      probs, action_samples, _ = mcmc.iterative_dfo(
          self._actor_network, batch_size, maybe_tiled_obs, action_samples, policy_state,
          num_action_samples=n, min_actions=mn, max_actions=mx, training=self._training,
          late_fusion=self._late_fusion, tfa_step_type=time_step.step_type,
          uniform_draws=draws.get('dfo_uniforms'), normal_draws=draws.get('dfo_normals'))
    if self._use_langevin:
      action_samples = mcmc.langevin_actions_given_obs(
          self._actor_network, maybe_tiled_obs, action_samples, policy_state=policy_state,
          num_action_samples=n, min_actions=mn, max_actions=mx, training=False,
          tfa_step_type=time_step.step_type, noise_scale=1.0,
          normal_draws=draws.get('langevin_normals'))
      if self._optimize_again:
        action_samples = mcmc.langevin_actions_given_obs(
            self._actor_network, maybe_tiled_obs, action_samples, policy_state=policy_state,
            num_action_samples=n, min_actions=mn, max_actions=mx, training=False,
            tfa_step_type=time_step.step_type,
            sampler_stepsize_init=self._again_stepsize_init,
            sampler_stepsize_final=self._again_stepsize_final,
            noise_scale=self._inference_langevin_noise_scale,
            normal_draws=draws.get('again_normals'))
      probs = mcmc.get_probabilities(self._actor_network, batch_size, n, maybe_tiled_obs,
                                     action_samples, training=False)
    if self._act_denorm_layer is not None:
      action_samples = self._act_denorm_layer(action_samples)
    distribution = MappedCategorical(probs=probs, mapped_values=action_samples)
    distribution.batch_size = batch_size
    return specs.PolicyStep(distribution, policy_state, ())

  def distribution(self, time_step, policy_state=(), **kwargs):
    return self._distribution(time_step, policy_state, **kwargs)

  def action(self, time_step, policy_state=(), seed=None, **kwargs):
    """Samples from the distribution (collect-policy behaviour)."""
    step = self._distribution(time_step, policy_state, **kwargs)
    action = step.action.sample(seed=seed)
    return specs.PolicyStep(self._clip_to_spec(action), step.state, step.info)

  def _clip_to_spec(self, action):
    if not self._clip or not hasattr(self.action_spec, 'minimum'):
      return action
    bounds = getattr(self, '_clip_bounds', None)
    if bounds is None or bounds[0].device != action.device:
      # device copies made once: a per-call host -> device copy is a synchronous transfer
      # that waits for the sampler's kernels (it was half of an action's wall time)
      bounds = tuple(torch.as_tensor(np.asarray(v, np.float32)).to(action.device).reshape(-1)
                     for v in (self.action_spec.minimum, self.action_spec.maximum))
      self._clip_bounds = bounds
    return torch.minimum(torch.maximum(action, bounds[0]), bounds[1])


class GreedyPolicy:
  """tf-agents GreedyPolicy over an IbcPolicy: mode() of the distribution,
  taken per observation (== the reference's flat arg max at B=1)."""

  def __init__(self, policy: IbcPolicy):
    self._wrapped_policy = policy
    self.time_step_spec = policy.time_step_spec
    self.action_spec = policy.action_spec

  @property
  def wrapped_policy(self):
    return self._wrapped_policy

  def get_initial_state(self, batch_size=None):
    return self._wrapped_policy.get_initial_state(batch_size)

  def distribution(self, time_step, policy_state=(), **kwargs):
    return self._wrapped_policy.distribution(time_step, policy_state, **kwargs)

  def action(self, time_step, policy_state=(), seed=None, **kwargs):
    del seed
    step = self._wrapped_policy.distribution(time_step, policy_state, **kwargs)
    d = step.action
    b = d.batch_size
    probs = d.probs.reshape(b, -1)
    values = d.mapped_values.reshape(b, probs.shape[1], -1)
    wp = self._wrapped_policy
    if probs.is_cuda and probs.dtype == torch.float32 and values.dtype == torch.float32:
      # arg max, gather and clip in one launch (ibc_greedy_actions)
      bounds = (None, None)
      if wp._clip and hasattr(wp.action_spec, 'minimum'):
        bounds = getattr(self, '_greedy_bounds', None)
        if bounds is None or bounds[0].device != values.device:
          wp._clip_to_spec(values[:1, 0])                    # creates the cached device bounds
          a_dim = values.shape[-1]
          bounds = tuple(v.expand(a_dim).contiguous() for v in wp._clip_bounds)
          self._greedy_bounds = bounds
      action = eng.greedy_actions(probs.contiguous(), values.contiguous(), *bounds)
      return specs.PolicyStep(action, step.state, step.info)
    idx = torch.argmax(probs, dim=1)
    action = values[torch.arange(b, device=values.device), idx]
    return specs.PolicyStep(wp._clip_to_spec(action), step.state, step.info)
