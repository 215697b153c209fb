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


# ---- PixelEBM on the general tf32 GEMM (csrc/gemm_tc.cu) ------------------------------
# Every operand of a contraction is rounded to nearest on the way into shared memory,
# in the forward, data-gradient and weight-gradient products alike.


class _RoundOperand(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x):
    return round_tf32(x)

  @staticmethod
  def backward(ctx, g):
    return g


class _RoundGradient(torch.autograd.Function):
  @staticmethod
  def forward(ctx, y):
    return y.clone()

  @staticmethod
  def backward(ctx, g):
    return round_tf32(g)


def _mm(a, w):
  return _RoundGradient.apply(_RoundOperand.apply(a) @ _RoundOperand.apply(w))


def pixel_energy(flat: torch.Tensor, spec, images: torch.Tensor, act: torch.Tensor) -> torch.Tensor:
  """Same signature as oracle.pixel_ebm.energy (networks/pixel_ebm.py:98-125) with tf32
  operands in the convolutions and the value head; the K = act_dim product and the
  [B, 256] encoding projection stay exact (the CUDA path keeps blocks of fewer than 64
  rows or K < 16 on the fp32 kernel)."""
  import torch.nn.functional as F
  from oracle import pixel_ebm as opix
  p = opix.unpack(flat, spec)
  x = opix.preprocess(images, spec).permute(0, 3, 1, 2)
  for i in range(4):
    k = p[f'conv{i + 1}_k'].permute(3, 2, 0, 1)
    y = F.conv2d(_RoundOperand.apply(x), _RoundOperand.apply(k), None, padding=1)
    x = F.relu(_RoundGradient.apply(y) + p[f'conv{i + 1}_b'].reshape(1, -1, 1, 1))
    x = F.max_pool2d(x, 2)
  enc = x.mean(dim=(2, 3))
  return pixel_value(flat, spec, enc, act)


def pixel_value(flat: torch.Tensor, spec, enc: torch.Tensor, act: torch.Tensor) -> torch.Tensor:
  """The value head of pixel_energy on an un-tiled encoding [B, 256] and act [B*n, A]
  (PixelEBM.call with observation_encoding, pixel_ebm.py:113-125)."""
  from oracle import pixel_ebm as opix
  p = opix.unpack(flat, spec)
  n = act.shape[0] // enc.shape[0]
  w0 = p['v_dense0_W']
  x = torch.repeat_interleave(enc @ w0[:256] + p['v_dense0_b'], n, dim=0) + act @ w0[256:]
  for j in range(spec.value_blocks):
    y = _mm(torch.relu(x), p[f'v_blk{j}_d0_W']) + p[f'v_blk{j}_d0_b']
    y = _mm(torch.relu(y), p[f'v_blk{j}_d1_W']) + p[f'v_blk{j}_d1_b']
    y = _mm(torch.relu(y), p[f'v_blk{j}_d2_W']) + p[f'v_blk{j}_d2_b']
    x = x + y
  return x @ p['v_dense1_W'] + p['v_dense1_b']
