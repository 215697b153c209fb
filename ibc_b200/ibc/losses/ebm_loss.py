"""EBM losses of the reference `ibc/losses/ebm_loss.py`, on the B200 engine.
`info_nce` (ebm_loss.py:21-48) is the hot path; `cd` and `clipped_cd` (:51-66, :109-148; selected by
no shipped gin) run on the same engine (`ibc_cd_loss`, and inside
`ibc_train_grads` through `ibc_loss_cfg.ebm_loss_type`); `cd_kl` is not built."""
from __future__ import annotations

import torch

from ibc_b200 import engine as eng


def info_nce(predictions,
             batch_size,
             num_counter_examples,
             softmax_temperature,
             kl=None):
  """ebm_loss.py:21-48.  predictions [B, n+1] with the true action LAST.
  Returns (per_example_loss [B], dict()).  `kl` (a Keras KLDivergence object
  upstream) is accepted and ignored: the kernel implements
  KLDivergence(reduction=NONE) with its epsilon clipping."""
  del kl
  predictions = predictions.to(torch.float32).contiguous()
  if tuple(predictions.shape) != (int(batch_size), int(num_counter_examples) + 1):
    raise ValueError('predictions must be [batch_size, num_counter_examples + 1], got %s'
                     % (tuple(predictions.shape),))
  per_example_loss = eng.info_nce(predictions, float(softmax_temperature))
  return per_example_loss, dict()


def _combined(counter_example_actions, true_actions):
  """[B, n, A] counter examples + [B, 1 | n, A] true actions -> [B * (n + 1), A], true LAST."""
  true = true_actions[:, :1, :]
  return torch.cat([counter_example_actions, true], dim=1).reshape(
      -1, counter_example_actions.shape[-1]).to(torch.float32).contiguous()


def cd(predictions):
  """ebm_loss.py:51-66: mean of the counter examples' predictions minus the true one."""
  predictions = predictions.to(torch.float32).contiguous()
  return eng.cd_loss(predictions, None, 'cd'), dict()


def cd_kl(*unused_args, **unused_kwargs):
  raise NotImplementedError("ebm_loss_type 'cd_kl' is out of scope (no shipped gin selects it; it "
                            "needs a copy of the network and gradients through the sampler)")


def clipped_cd(predictions, counter_example_actions, actions_size_n, soft):
  """ebm_loss.py:109-148.  counter_example_actions, actions_size_n: [B, n, A]."""
  predictions = predictions.to(torch.float32).contiguous()
  combined = _combined(counter_example_actions, actions_size_n)
  kind = 'soft_clipped_cd' if soft else 'clipped_cd'
  per_example_loss = eng.cd_loss(predictions, combined, kind)
  return per_example_loss, clipped_cd_debug(predictions, counter_example_actions, actions_size_n, soft)


def clipped_cd_debug(predictions, counter_example_actions, actions_size_n, soft):
  """The logged scalars of clipped_cd (ebm_loss.py:138-147), off the hot path."""
  errors = torch.linalg.norm(counter_example_actions - actions_size_n, dim=-1) ** 2
  outside = (errors > 2e-1).to(predictions.dtype)
  debug = {}
  if soft:
    data, samp = predictions[:, -1:], predictions[:, :-1]
    debug['cd_loss'] = ((samp - data) * outside).mean(dim=1).mean()
    debug['soft_loss'] = (((data - samp) - errors).abs() * (1.0 - outside)).mean(dim=1).mean()
  debug['fraction_outside_threshold'] = outside.mean()
  return debug
