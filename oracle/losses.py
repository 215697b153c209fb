"""Oracle (test infrastructure): ibc/losses/ebm_loss.py:21-48 (info_nce) and
ibc/losses/gradient_loss.py:23-72 (grad_penalty), in PyTorch-CPU.

Third-party arithmetic restated: Keras 2.6.0 `KLDivergence(reduction=NONE)`
(keras/losses.py kl_divergence): y_true and y_pred are both clipped to
[epsilon, 1] with epsilon = 1e-7 BEFORE sum(y_true * log(y_true / y_pred)).
With one-hot labels the zero entries become 1e-7, so this is not plain
cross-entropy.
"""
from __future__ import annotations

import torch

from oracle import mcmc as omcmc

KERAS_EPSILON = 1e-7


def keras_kl_divergence(y_true: torch.Tensor, y_pred: torch.Tensor):
  y_true = torch.clamp(y_true, KERAS_EPSILON, 1.0)
  y_pred = torch.clamp(y_pred, KERAS_EPSILON, 1.0)
  return torch.sum(y_true * torch.log(y_true / y_pred), dim=-1)


def info_nce(predictions: torch.Tensor, batch_size: int,
             num_counter_examples: int, softmax_temperature: float = 1.0):
  """predictions [B, n+1], true action in the LAST column -> loss [B]."""
  assert predictions.shape == (batch_size, num_counter_examples + 1)
  p = torch.softmax(predictions / softmax_temperature, dim=-1)
  labels = torch.zeros_like(p)
  labels[:, num_counter_examples] = 1.0
  return keras_kl_divergence(labels, p)


def cd(predictions: torch.Tensor):
  """ibc/losses/ebm_loss.py:51-66: mean over the counter examples minus the true action's."""
  return predictions[:, :-1].mean(dim=1) - predictions[:, -1:].mean(dim=1)


def clipped_cd(predictions: torch.Tensor, counter_example_actions: torch.Tensor,
               actions_size_n: torch.Tensor, soft: bool):
  """ibc/losses/ebm_loss.py:109-148 (thresh 2e-1, w_shaping 1.0).  Returns (loss [B], debug)."""
  energy_data = predictions[:, -1:].expand(-1, predictions.shape[1] - 1)
  energy_samp = predictions[:, :-1]
  errors = torch.linalg.norm(counter_example_actions - actions_size_n, dim=-1) ** 2
  outside = (errors > 2e-1).to(predictions.dtype)
  loss = ((energy_samp - energy_data) * outside).mean(dim=1)
  debug = {}
  if soft:
    inside = (errors <= 2e-1).to(predictions.dtype)
    residual = (energy_data - energy_samp) - errors
    soft_loss = 1.0 * (residual.abs() * inside).mean(dim=1)
    debug['cd_loss'] = loss.mean()
    debug['soft_loss'] = soft_loss.mean()
    loss = loss + soft_loss
  debug['fraction_outside_threshold'] = outside.mean()
  return loss, debug


def grad_penalty(energy_fn, grad_norm_type, batch_size, observations,
                 combined_true_counter_actions, grad_margin=1.0,
                 square_grad_penalty=True, grad_loss_weight=1.0,
                 create_graph=True):
  """only_apply_final_grad_penalty=True branch (gradient_loss.py:37-48,
  62-72); differentiable w.r.t. the parameters closed over by energy_fn."""
  a = combined_true_counter_actions.detach().clone().requires_grad_(True)
  e = energy_fn(observations, a)
  (g,) = torch.autograd.grad(e.sum(), a, create_graph=create_graph)
  de_dact = -g
  grad_norms = omcmc.compute_grad_norm(grad_norm_type, de_dact)
  grad_norms = grad_norms.reshape(batch_size, -1)
  if grad_margin is not None:
    grad_norms = torch.clamp(grad_norms - grad_margin, 0.0, 1e10)
  if square_grad_penalty:
    grad_norms = grad_norms ** 2
  return grad_norms.mean(dim=1) * grad_loss_weight
