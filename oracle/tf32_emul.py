"""Oracle (test infrastructure): the MLPEBM with every tensor-core operand
rounded to tf32 (10-bit mantissa, round-to-nearest, ties away), fp64
accumulation.  It is NOT a restatement of the reference (which is fp32): it
calibrates how far a correct kind::tf32 implementation of
networks/layers/resnet.py:135-153 may drift from the exact result, so that the
GPU tolerances for deep ReLU stacks (mask flips near zero, cancelling InfoNCE
gradients) are measured instead of guessed ("calibrate")."""
from __future__ import annotations

import torch

from oracle import mlp_ebm as omlp


def round_tf32(x: torch.Tensor) -> torch.Tensor:
  xi = x.float().contiguous().view(torch.int32)
  r = (xi + 0x1000) & ~0x1FFF
  return r.view(torch.float32).to(x.dtype)


class _MatmulTF32(torch.autograd.Function):
  @staticmethod
  def forward(ctx, a, w):
    ar, wr = round_tf32(a), round_tf32(w)
    ctx.save_for_backward(ar, wr)
    return ar @ wr

  @staticmethod
  def backward(ctx, g):
    ar, wr = ctx.saved_tensors
    # dgrad operand rounded to nearest; wgrad operand truncated by the MMA
    gi = g.float().contiguous().view(torch.int32) & ~0x1FFF
    gt = gi.view(torch.float32).to(g.dtype)
    return round_tf32(g) @ wr.t(), ar.t() @ gt


def energy(flat: torch.Tensor, spec: omlp.MLPSpec, obs: torch.Tensor,
           act: torch.Tensor) -> torch.Tensor:
  """Same signature as oracle.mlp_ebm.energy; the observation half of the
  first layer is exact (the CUDA path hoists it in fp32)."""
  p = omlp.unpack(flat, spec)
  o = spec.obs_dim
  h = obs @ p['Wp'][:o] + p['bp'] + _MatmulTF32.apply(act, p['Wp'][o:])
  for j in range(spec.num_blocks):
    v = _MatmulTF32.apply(torch.relu(h), p[f'W1_{j}']) + p[f'b1_{j}']
    h = h + _MatmulTF32.apply(torch.relu(v), p[f'W2_{j}']) + p[f'b2_{j}']
  return h @ p['we'] + p['be']
