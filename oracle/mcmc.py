"""Oracle (test infrastructure): samplers of ibc/agents/mcmc.py.

Every random draw is an explicit input so the CUDA path can be checked on
identical numbers:
  * DFO resampling takes `uniforms` (fp64 in [0, 1), one per draw) and, for the
    perturbation, `normals` (standard normal, fp32);
  * Langevin takes `normals` [K, R, A].

`energy_fn(obs, actions) -> E[R]` is any torch callable (an oracle MLPEBM
closure, or the mock quadratic energy of mcmc_test.py:67-81).

Third-party arithmetic restated here (TensorFlow 2.6.0,
tensorflow/core/kernels/multinomial_op.cc, MultinomialFunctor<CPUDevice>):
per row, max over finite logits; cdf[j] = sum_{k<=j} exp(double(logit_k) -
double(max)) accumulated SEQUENTIALLY in fp64, non-finite logits skipped;
each draw: to_find = u * total; index = upper_bound(cdf, to_find).
Deviations (documented, measure-zero): exp is the portable `exp64` below
(same IEEE operation sequence in NumPy, C and CUDA, so all three agree
bit-for-bit; TF uses Eigen's pexp), arguments below -708 flush to 0, and an
index that would land at n (u*total rounding up to total) is clamped to n-1.
"""
from __future__ import annotations

import collections
import math
from typing import Callable, Optional

import numpy as np
import torch

# ---------------------------------------------------------------------------
# Portable fp64 exp: only mul/add/rint on IEEE doubles, no FMA, so that NumPy,
# C (-ffp-contract=off) and CUDA (__dmul_rn/__dadd_rn) give identical bits.
# ---------------------------------------------------------------------------
EXP64_LOG2E = 1.4426950408889634
EXP64_LN2_HI = 6.93147180369123816490e-01   # low 32 mantissa bits are zero
EXP64_LN2_LO = 1.90821492927058770002e-10
EXP64_MIN_ARG = -708.0
EXP64_COEF = [1.0 / math.factorial(i) for i in range(14)]   # Taylor, degree 13


def exp64(x: np.ndarray) -> np.ndarray:
  """exp(x) for x <= 0 (fp64), flushing x < -708 to 0."""
  x = np.asarray(x, dtype=np.float64)
  xs = np.where(x < EXP64_MIN_ARG, 0.0, x)
  k = np.rint(xs * EXP64_LOG2E)
  r = (xs - k * EXP64_LN2_HI) - k * EXP64_LN2_LO
  p = np.full_like(r, EXP64_COEF[13])
  for i in range(12, -1, -1):
    p = p * r + EXP64_COEF[i]
  out = p * np.ldexp(1.0, k.astype(np.int64))
  return np.where(x < EXP64_MIN_ARG, 0.0, out)


def multinomial_cdf(log_ps_row: np.ndarray):
  """(cdf[n] fp64, total) of one row, TF CPU multinomial semantics."""
  logits = np.asarray(log_ps_row, dtype=np.float32)
  finite = np.isfinite(logits)
  if finite.any():
    mx = np.float64(logits[finite].max())
  else:
    mx = np.float64(np.finfo(np.float32).min)   # numeric_limits<T>::lowest()
  w = exp64(np.where(finite, logits.astype(np.float64) - mx, 0.0))
  w = np.where(finite, w, 0.0)
  cdf = np.add.accumulate(w)      # strictly sequential fp64 accumulation
  return cdf, (cdf[-1] if len(cdf) else np.float64(0.0))


def multinomial_draws(log_ps: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
  """Raw drawn class indices [B, count] (tf.random.categorical, mcmc.py:51)."""
  log_ps = np.asarray(log_ps, dtype=np.float32)
  uniforms = np.asarray(uniforms, dtype=np.float64)
  b, n = log_ps.shape
  out = np.zeros(uniforms.shape, dtype=np.int32)
  for i in range(b):
    cdf, total = multinomial_cdf(log_ps[i])
    to_find = uniforms[i] * total
    idx = np.searchsorted(cdf, to_find, side='right')     # std::upper_bound
    out[i] = np.minimum(idx, n - 1)
  return out


def categorical_bincount(count: int, log_ps, n: int,
                         uniforms: np.ndarray) -> np.ndarray:
  """mcmc.py:32-55: counts [B, n] int32 summing to `count` per row."""
  log_ps = np.asarray(log_ps, dtype=np.float32)
  assert log_ps.shape[1] == n and uniforms.shape == (log_ps.shape[0], count)
  draws = multinomial_draws(log_ps, uniforms)
  counts = np.zeros((log_ps.shape[0], n), dtype=np.int32)
  for i in range(log_ps.shape[0]):
    counts[i] = np.bincount(draws[i], minlength=n)[:n]
  return counts


def resample_indices(log_ps, uniforms) -> np.ndarray:
  """Global row ids after tf.repeat(range(B*n), counts) (mcmc.py:136-137):
  per observation the drawn indices SORTED ascending, offset by b*n."""
  log_ps = np.asarray(log_ps, dtype=np.float32)
  b, n = log_ps.shape
  counts = categorical_bincount(uniforms.shape[1], log_ps, n, uniforms)
  return np.repeat(np.arange(b * n, dtype=np.int64), counts.reshape(-1))


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
  return torch.softmax(x, dim=1)


def iterative_dfo(energy_fn: Callable, batch_size: int, observations,
                  action_samples: torch.Tensor, num_action_samples: int,
                  min_actions: torch.Tensor, max_actions: torch.Tensor,
                  uniforms: np.ndarray, normals: Optional[torch.Tensor],
                  temperature: float = 1.0, num_iterations: int = 3,
                  iteration_std: float = 0.33, return_debug: bool = False):
  """mcmc.py:58-158.  uniforms [num_iterations, B, n] fp64; normals
  [num_iterations-1, B*n, A].  Returns (probs[B*n], actions[B*n, A])."""
  n = num_action_samples
  dbg = {'logits': [], 'indices': []}

  def update_selected_actions(samples, it):
    net_logits = energy_fn(observations, samples)              # :112-116
    log_probs = net_logits.reshape(batch_size, n) / temperature  # :121-125
    idx = resample_indices(log_probs.detach().numpy(), uniforms[it])
    assert idx.shape[0] == batch_size * n
    dbg['logits'].append(log_probs.detach().clone())
    dbg['indices'].append(idx.copy())
    return log_probs, samples[torch.from_numpy(idx)]            # :139-140

  samples = action_samples
  std = iteration_std
  log_probs, samples = update_selected_actions(samples, 0)       # :142-143
  for it in range(1, num_iterations):                            # :145
    samples = samples + normals[it - 1].to(samples.dtype) * std  # :146-147
    samples = torch.minimum(torch.maximum(samples, min_actions), max_actions)
    log_probs, samples = update_selected_actions(samples, it)
    std *= 0.5                                                   # :153
  probs = softmax_rows(log_probs).reshape(-1)                    # :155-156
  if return_debug:
    return probs, samples, dbg
  return probs, samples


def gradient_wrt_act(energy_fn, observations, actions, apply_exp=False):
  """mcmc.py:161-191: (-dE/dact, E)."""
  a = actions.detach().clone().requires_grad_(True)
  e = energy_fn(observations, a)
  e_used = torch.exp(e) if apply_exp else e
  (g,) = torch.autograd.grad(e_used.sum(), a)
  return -g, e_used.detach()


def compute_grad_norm(grad_norm_type, de_dact: torch.Tensor) -> torch.Tensor:
  """mcmc.py:194-206."""
  if grad_norm_type is None:
    return torch.zeros_like(de_dact[:, 0])
  ord_ = {'1': 1, '2': 2, 'inf': float('inf')}[grad_norm_type]
  return torch.linalg.norm(de_dact, ord=ord_, dim=1)


class PolynomialSchedule:
  """mcmc.py:281-294."""

  def __init__(self, init, final, power, num_steps):
    self._init, self._final = init, final
    self._power, self._num_steps = power, num_steps

  def get_rate(self, index):
    return ((self._init - self._final) *
            ((1 - (float(index) / float(self._num_steps - 1))) ** self._power)
            ) + self._final


class ExponentialSchedule:
  """mcmc.py:267-278."""

  def __init__(self, init, decay):
    self._decay, self._latest_lr = decay, init

  def get_rate(self, index):
    del index
    self._latest_lr *= self._decay
    return self._latest_lr


def stepsize_table(num_iterations, init=1e-1, decay=0.8, use_polynomial=True,
                   final=1e-5, power=2.0):
  """Step size used AT step i (step 0 uses init; step i>=1 uses
  get_rate(i), because the reference asks for `step_index + 1` after each
  step, mcmc.py:359,408)."""
  sched = (PolynomialSchedule(init, final, power, num_iterations)
           if use_polynomial else ExponentialSchedule(init, decay))
  out = [init]
  for i in range(num_iterations - 1):
    out.append(sched.get_rate(i + 1))
  return out


ChainData = collections.namedtuple('ChainData',
                                   ['actions', 'energies', 'grad_norms'])


def langevin_step(energy_fn, observations, actions, noise, noise_scale,
                  grad_clip, delta_action_clip, stepsize, min_actions,
                  max_actions, grad_norm_type, apply_exp=False):
  """mcmc.py:209-264 with the tf.random.normal draw supplied as `noise`."""
  de_dact, energies = gradient_wrt_act(energy_fn, observations, actions,
                                       apply_exp)
  dclip = delta_action_clip * 0.5 * (max_actions - min_actions)
  grad_norms = compute_grad_norm(grad_norm_type, de_dact)
  if grad_clip is not None:
    de_dact = torch.clamp(de_dact, -grad_clip, grad_clip)
  de_dact = 0.5 * de_dact + noise.to(de_dact.dtype) * noise_scale
  delta = stepsize * de_dact
  delta = torch.minimum(torch.maximum(delta, -dclip), dclip)
  actions = actions - delta
  actions = torch.minimum(torch.maximum(actions, min_actions), max_actions)
  return actions, energies, grad_norms


def langevin_actions_given_obs(energy_fn, observations, action_samples,
                               min_actions, max_actions, normals,
                               num_iterations=25, sampler_stepsize_init=1e-1,
                               sampler_stepsize_decay=0.8, noise_scale=1.0,
                               grad_clip=None, delta_action_clip=0.1,
                               apply_exp=False, use_polynomial_rate=True,
                               sampler_stepsize_final=1e-5,
                               sampler_stepsize_power=2.0, return_chain=False,
                               grad_norm_type='inf'):
  """mcmc.py:330-425.  normals [num_iterations, R, A]."""
  steps = stepsize_table(num_iterations, sampler_stepsize_init,
                         sampler_stepsize_decay, use_polynomial_rate,
                         sampler_stepsize_final, sampler_stepsize_power)
  actions = action_samples.clone()
  chain_a, chain_e, chain_g = [], [], []
  for i in range(num_iterations):
    actions, energies, grad_norms = langevin_step(
        energy_fn, observations, actions, normals[i], noise_scale, grad_clip,
        delta_action_clip, steps[i], min_actions, max_actions, grad_norm_type,
        apply_exp)
    actions = actions.detach()
    if return_chain:
      chain_a.append(actions.clone())       # post-step actions  (:297-327)
      chain_e.append(energies.clone())      # pre-step energies
      chain_g.append(grad_norms.clone())    # pre-step grad norms
  if return_chain:
    return actions, ChainData(torch.stack(chain_a), torch.stack(chain_e),
                              torch.stack(chain_g))
  return actions


def get_probabilities(energy_fn, batch_size, num_action_samples, observations,
                      actions, temperature=1.0):
  """mcmc.py:428-441."""
  logits = energy_fn(observations, actions).reshape(batch_size,
                                                   num_action_samples)
  return softmax_rows(logits / temperature).reshape(-1)
