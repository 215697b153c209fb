"""EBM losses of the reference `ibc/losses/ebm_loss.py`, on the B200 engine.
Only `info_nce` (ebm_loss.py:21-48) is on the hot path; `cd`, `cd_kl` and
`clipped_cd` are selected by no shipped gin."""
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


def cd(*unused_args, **unused_kwargs):
  raise NotImplementedError("ebm_loss_type 'cd' is out of scope (no shipped gin selects it)")


def cd_kl(*unused_args, **unused_kwargs):
  raise NotImplementedError("ebm_loss_type 'cd_kl' is out of scope (no shipped gin selects it)")


def clipped_cd(*unused_args, **unused_kwargs):
  raise NotImplementedError("ebm_loss_type 'clipped_cd' is out of scope (no shipped gin selects it)")
