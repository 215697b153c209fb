"""Gradient penalty of the reference `ibc/losses/gradient_loss.py:23-72`."""
from __future__ import annotations

import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin
from ibc_b200.ibc.agents import mcmc


@gin.configurable
def grad_penalty(energy_network,
                 grad_norm_type,
                 batch_size,
                 chain_data,
                 observations,
                 combined_true_counter_actions,
                 training,
                 only_apply_final_grad_penalty=True,
                 grad_margin=1.0,
                 square_grad_penalty=True,
                 grad_loss_weight=1.0):
  """Forward value of the penalty, [B] (the training gradient of it is
  produced inside `ibc_train_grads`, see ImplicitBCAgent)."""
  cfg = eng.loss_cfg(1.0, True, grad_norm_type, grad_margin, square_grad_penalty,
                     grad_loss_weight)
  if only_apply_final_grad_penalty:
    de_dact, _ = mcmc.gradient_wrt_act(energy_network, observations,
                                       combined_true_counter_actions.detach(), training,
                                       network_state=(), tfa_step_type=(), apply_exp=False)
    return eng.grad_penalty_from_dact(de_dact.contiguous(), int(batch_size), cfg)
  # Case 2 (gradient_loss.py:51-60): penalise every norm of the chain.
  assert chain_data.grad_norms is not None
  grad_norms = chain_data.grad_norms.transpose(0, 1).reshape(int(batch_size), -1)
  if grad_margin is not None:
    grad_norms = torch.clamp(grad_norms - grad_margin, 0.0, 1e10)
  if square_grad_penalty:
    grad_norms = grad_norms ** 2
  return grad_norms.mean(dim=1) * grad_loss_weight
