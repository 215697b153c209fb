"""Oracle (test infrastructure): the agent/policy level of the hot path.

Restates, in PyTorch-CPU with explicit random draws:
  ibc/agents/ibc_agent.py:340-475  _make_counter_example_actions
  ibc/agents/ibc_agent.py:132-300  _loss (info_nce + grad_penalty, metrics)
  ibc/agents/base_agent.py:36-62   _train (loss -> gradients -> Adam)
  ibc/train/get_agent.py:64-68     Adam + ExponentialDecay(decay_rate=0.99)
  ibc/agents/ibc_policy.py:240-328 IbcPolicy._distribution + greedy mode
Third-party arithmetic restated: tf-agents `common.aggregate_losses`
(sum(per_example) / (local_batch * num_replicas)), Keras 2.6 Adam
(theta -= lr_t * m / (sqrt(v) + eps), lr_t = lr * sqrt(1-b2^t) / (1-b1^t),
eps = 1e-7), Keras ExponentialDecay (lr * rate ** (step / decay_steps),
continuous).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, Optional

import numpy as np
import torch

from oracle import losses as olosses
from oracle import mcmc as omcmc


def tile_batch(x: torch.Tensor, multiplier: int) -> torch.Tensor:
  """tf-agents nest_utils.tile_batch: each entry repeated contiguously."""
  return torch.repeat_interleave(x, multiplier, dim=0)


def uniform_actions(u: torch.Tensor, min_actions, max_actions):
  """tensor_spec.sample_spec_nest on a float BoundedTensorSpec: u*(max-min)+min
  (ibc_agent.py:350-352, ibc_policy.py:263-265)."""
  return u * (max_actions - min_actions) + min_actions


@dataclasses.dataclass
class AgentConfig:
  num_counter_examples: int = 256
  fraction_dfo_samples: float = 0.0
  fraction_langevin_samples: float = 1.0
  add_grad_penalty: bool = True
  grad_norm_type: Optional[str] = 'inf'
  softmax_temperature: float = 1.0
  compute_mse: bool = False
  # langevin_actions_given_obs gin bindings
  langevin_num_iterations: int = 25
  langevin_stepsize_init: float = 1e-1
  langevin_stepsize_final: float = 1e-5
  langevin_stepsize_power: float = 2.0
  langevin_noise_scale: float = 1.0
  langevin_delta_action_clip: float = 0.1
  # grad_penalty gin bindings
  grad_margin: float = 1.0
  grad_loss_weight: float = 1.0


def make_counter_example_actions(energy_fn, cfg: AgentConfig, obs, true_actions,
                                 min_actions, max_actions, uniform_u,
                                 langevin_normals=None, dfo_uniforms=None,
                                 dfo_normals=None):
  """ibc_agent.py:340-475.  obs [B, O]; true_actions [B, A]; uniform_u
  [B, N, A] in [0,1).  Returns (counter [B,N,A], combined [B*(N+1),A], chain).
  """
  b, n = obs.shape[0], cfg.num_counter_examples
  rnd = uniform_actions(uniform_u, min_actions, max_actions)     # [B, N, A]
  chain = None
  if cfg.fraction_dfo_samples == 0.0 and cfg.fraction_langevin_samples == 0.0:
    counter = rnd
  else:
    flat = rnd.reshape(b * n, -1)
    obs_n = tile_batch(obs, n)
    dfo_actions = lang_actions = None
    if cfg.fraction_dfo_samples > 0.0:
      _, dfo_actions = omcmc.iterative_dfo(
          energy_fn, b, obs_n, flat, n, min_actions, max_actions,
          dfo_uniforms, dfo_normals)
    if cfg.fraction_langevin_samples > 0.0:
      lang_actions, chain = omcmc.langevin_actions_given_obs(
          energy_fn, obs_n, flat, min_actions, max_actions, langevin_normals,
          num_iterations=cfg.langevin_num_iterations,
          sampler_stepsize_init=cfg.langevin_stepsize_init,
          sampler_stepsize_final=cfg.langevin_stepsize_final,
          sampler_stepsize_power=cfg.langevin_stepsize_power,
          noise_scale=cfg.langevin_noise_scale,
          delta_action_clip=cfg.langevin_delta_action_clip,
          return_chain=True, grad_norm_type=cfg.grad_norm_type)
    frac_init = 1.0 - cfg.fraction_dfo_samples - cfg.fraction_langevin_samples
    n_init = int(frac_init * n)
    n_dfo = int(cfg.fraction_dfo_samples * n)
    n_lang = int(cfg.fraction_langevin_samples * n)
    residual = n - n_init - n_dfo - n_lang
    assert residual >= 0
    n_init += residual
    pieces, used = [], 0
    if n_init > 0:
      pieces.append(flat.reshape(b, n, -1)[:, :n_init])
      used += n_init
    if n_dfo > 0:
      pieces.append(dfo_actions.reshape(b, n, -1)[:, used:used + n_dfo])
      used += n_dfo
    if n_lang > 0:
      pieces.append(lang_actions.reshape(b, n, -1)[:, used:used + n_lang])
      used += n_lang
    assert used == n
    counter = torch.cat(pieces, dim=1)
  combined = torch.cat([counter, true_actions[:, None, :]], dim=1)
  combined = combined.reshape(b * (n + 1), -1)      # true action LAST per obs
  return counter, combined, chain


def loss(energy_fn, cfg: AgentConfig, obs, true_actions, min_actions,
         max_actions, uniform_u, global_batch: Optional[int] = None, **draws):
  """ibc_agent.py:132-300 (training=True forward).  Returns
  (total_loss scalar, per_example [B], metrics dict)."""
  b, n = obs.shape[0], cfg.num_counter_examples
  # The samplers detach internally (stop_chain_grad=True, and the loss takes
  # stop_gradient(actions), ibc_agent.py:210), so no gradient reaches theta
  # through the chain even though it runs "under the tape".
  counter, combined, chain = make_counter_example_actions(
      energy_fn, cfg, obs, true_actions, min_actions, max_actions, uniform_u,
      **draws)
  counter, combined = counter.detach(), combined.detach()
  obs_t = tile_batch(obs, n + 1)
  predictions = energy_fn(obs_t, combined.detach()).reshape(b, n + 1)
  per_example = olosses.info_nce(predictions, b, n, cfg.softmax_temperature)
  metrics: Dict[str, torch.Tensor] = {}
  if cfg.add_grad_penalty:
    gl = olosses.grad_penalty(energy_fn, cfg.grad_norm_type, b, obs_t,
                              combined, cfg.grad_margin,
                              grad_loss_weight=cfg.grad_loss_weight)
    per_example = per_example + gl
    metrics['grad_loss'] = gl.mean().detach()
  denom = float(global_batch if global_batch is not None else b)
  total = per_example.sum() / denom              # common.aggregate_losses
  metrics['ebm_total_loss'] = total.detach()
  if cfg.compute_mse:
    # Keras MeanSquaredError(NONE) over the action axis -> [B, N]; tf-agents
    # aggregate_losses averages the extra axes of a rank>1 loss before the
    # sum / global_batch (ibc_agent.py:182-189).  Logging only.
    mse = ((counter - true_actions[:, None, :]) ** 2).mean(dim=-1).mean(dim=1)
    metrics['mse_counter_examples'] = (mse.sum() / denom).detach()
  if chain is not None:
    metrics['overall_energies_avg'] = chain.energies.mean()
    metrics['first_energies_avg'] = chain.energies[0].mean()
    metrics['final_energies_avg'] = chain.energies[-1].mean()
    metrics['overall_grad_norms_avg'] = chain.grad_norms.mean()
    metrics['first_grad_norms_avg'] = chain.grad_norms[0].mean()
    metrics['final_grad_norms_avg'] = chain.grad_norms[-1].mean()
  return total, per_example, metrics


# ---------------------------------------------------------------------------
# Optimiser (ibc/train/get_agent.py:64-68; Keras 2.6 Adam, non-amsgrad).
# ---------------------------------------------------------------------------
@dataclasses.dataclass
class AdamState:
  m: torch.Tensor
  v: torch.Tensor
  iterations: int = 0


def adam_init(flat: torch.Tensor) -> AdamState:
  return AdamState(torch.zeros_like(flat), torch.zeros_like(flat), 0)


def exponential_decay_lr(base_lr: float, step: int, decay_steps: int = 100,
                         decay_rate: float = 0.99) -> float:
  return base_lr * decay_rate ** (step / decay_steps)


def adam_step_scalars(base_lr, iterations, beta1=0.9, beta2=0.999,
                      decay_steps=100, decay_rate=0.99):
  """lr_t used by the step that turns `iterations` into `iterations + 1`."""
  t = iterations + 1
  lr = exponential_decay_lr(base_lr, iterations, decay_steps, decay_rate)
  return lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)


def adam_update(flat, grad, state: AdamState, base_lr, beta1=0.9, beta2=0.999,
                eps=1e-7, decay_steps=100, decay_rate=0.99):
  lr_t = adam_step_scalars(base_lr, state.iterations, beta1, beta2,
                           decay_steps, decay_rate)
  state.m = beta1 * state.m + (1.0 - beta1) * grad
  state.v = beta2 * state.v + (1.0 - beta2) * grad * grad
  new = flat - lr_t * state.m / (torch.sqrt(state.v) + eps)
  state.iterations += 1
  return new.to(flat.dtype)


def train_step(make_energy_fn: Callable, flat, state: AdamState,
               cfg: AgentConfig, obs, true_actions, min_actions, max_actions,
               uniform_u, base_lr, global_batch=None, **draws):
  """base_agent.py:36-62.  make_energy_fn(flat) -> energy_fn.  Returns
  (new_flat, total_loss, grad, metrics)."""
  theta = flat.detach().clone().requires_grad_(True)
  total, _, metrics = loss(make_energy_fn(theta), cfg, obs, true_actions,
                           min_actions, max_actions, uniform_u,
                           global_batch=global_batch, **draws)
  assert torch.isfinite(total), 'Loss is inf or nan'   # check_numerics, :46
  (grad,) = torch.autograd.grad(total, theta)
  new_flat = adam_update(flat.detach(), grad, state, base_lr)
  return new_flat, total.detach(), grad, metrics


# ---------------------------------------------------------------------------
# Policy (ibc/agents/ibc_policy.py:240-328 + GreedyPolicy mode()).
# ---------------------------------------------------------------------------
def policy_distribution(energy_fn, obs, num_action_samples, min_actions,
                        max_actions, uniform_u, use_dfo=False,
                        use_langevin=True, dfo_uniforms=None, dfo_normals=None,
                        langevin_normals=None, optimize_again=False,
                        again_normals=None, again_stepsize_init=1e-1,
                        again_stepsize_final=1e-5,
                        inference_langevin_noise_scale=1.0,
                        langevin_kwargs=None):
  """Returns (probs [B*n], actions [B*n, A]).  uniform_u [B*n, A]."""
  b, n = obs.shape[0], num_action_samples
  obs_t = tile_batch(obs, n)
  actions = uniform_actions(uniform_u, min_actions, max_actions)
  probs = torch.zeros(b * n)
  kw = dict(langevin_kwargs or {})
  if use_dfo:
    probs, actions = omcmc.iterative_dfo(
        energy_fn, b, obs_t, actions, n, min_actions, max_actions,
        dfo_uniforms, dfo_normals)
  if use_langevin:
    kw1 = dict(kw)
    kw1['noise_scale'] = 1.0                       # ibc_policy.py:294 explicit
    actions = omcmc.langevin_actions_given_obs(
        energy_fn, obs_t, actions, min_actions, max_actions, langevin_normals,
        **kw1)
    if optimize_again:
      kw2 = dict(kw)
      kw2.update(sampler_stepsize_init=again_stepsize_init,
                 sampler_stepsize_final=again_stepsize_final,
                 noise_scale=inference_langevin_noise_scale)
      actions = omcmc.langevin_actions_given_obs(
          energy_fn, obs_t, actions, min_actions, max_actions, again_normals,
          **kw2)
    with torch.no_grad():
      probs = omcmc.get_probabilities(energy_fn, b, n, obs_t, actions)
  return probs, actions


def greedy_action(probs, actions):
  """MappedCategorical.mode (ibc_policy.py:114-117): ONE argmax over the FLAT
  [B*n] probs -> action [1, A].  Only meaningful at B=1."""
  return actions[int(torch.argmax(probs))][None, :]


def greedy_action_per_observation(probs, actions, batch_size):
  """Batched extension used by the B200 build: argmax per observation; equals
  greedy_action at B=1."""
  n = probs.shape[0] // batch_size
  idx = torch.argmax(probs.reshape(batch_size, n), dim=1)
  return actions.reshape(batch_size, n, -1)[torch.arange(batch_size), idx]
