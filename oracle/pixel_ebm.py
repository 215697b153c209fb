"""Oracle (test infrastructure): late-fusion PixelEBM in PyTorch-CPU.

Restates:
  networks/utils/image_prepro.py:20-43     preprocess (u8->[0,1], RAW reshape
                                           [B,T,H,W,C]->[B,H,W,C*T], bilinear)
  networks/layers/conv_maxpool.py:20-48    4x(conv3x3 same + ReLU + maxpool2)+GAP
  networks/layers/dense_resnet_value.py:22-69  DenseResnetValue
  networks/pixel_ebm.py:79-125             encode / call (late fusion)
Third-party arithmetic restated: tf.image.resize bilinear in TF2 = half-pixel
centres, no antialias (== torch interpolate(mode='bilinear',
align_corners=False, antialias=False)); tf.image.convert_image_dtype u8->f32 =
x * (1/255); Keras Conv2D kernel layout [kh, kw, cin, cout], NHWC, 'same';
MaxPool2D() = 2x2 stride 2 'valid'; GlobalAveragePooling2D = mean over H, W.

Flat parameter layout (shared with the CUDA library):
  conv{1..4}: kernel [3,3,cin,cout] (Keras layout) then bias [cout],
     channels 6(=3T) -> 32 -> 64 -> 128 -> 256
  value dense0: W [256 + A, Wv], b [Wv]
  per block: d0 W [Wv, Wv/4], b; d1 W [Wv/4, Wv/4], b; d2 W [Wv/4, Wv], b
     (dense3 is never built: shapes always match, dense_resnet_value.py:67)
  value dense1: W [Wv] (Keras [Wv,1]), b [1]
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

CONV_CHANNELS = (32, 64, 128, 256)


@dataclasses.dataclass(frozen=True)
class PixelSpec:
  seq_len: int = 2
  height: int = 240
  width: int = 320
  channels: int = 3
  target_height: int = 180
  target_width: int = 240
  act_dim: int = 2
  value_width: int = 1024
  value_blocks: int = 1

  @property
  def in_channels(self) -> int:
    return self.channels * self.seq_len


def param_layout(spec: PixelSpec) -> List[Tuple[str, int, Tuple[int, ...]]]:
  out, off = [], 0

  def add(name, shape):
    nonlocal off
    out.append((name, off, tuple(shape)))
    off += math.prod(shape)

  cin = spec.in_channels
  for i, cout in enumerate(CONV_CHANNELS):
    add(f'conv{i + 1}_k', (3, 3, cin, cout))
    add(f'conv{i + 1}_b', (cout,))
    cin = cout
  w = spec.value_width
  add('v_dense0_W', (CONV_CHANNELS[-1] + spec.act_dim, w))
  add('v_dense0_b', (w,))
  for j in range(spec.value_blocks):
    add(f'v_blk{j}_d0_W', (w, w // 4)); add(f'v_blk{j}_d0_b', (w // 4,))
    add(f'v_blk{j}_d1_W', (w // 4, w // 4)); add(f'v_blk{j}_d1_b', (w // 4,))
    add(f'v_blk{j}_d2_W', (w // 4, w)); add(f'v_blk{j}_d2_b', (w,))
  add('v_dense1_W', (w,))
  add('v_dense1_b', (1,))
  return out


def param_count(spec: PixelSpec) -> int:
  name, off, shape = param_layout(spec)[-1]
  return off + math.prod(shape)


def unpack(flat, spec: PixelSpec) -> Dict[str, torch.Tensor]:
  return {name: flat[off:off + math.prod(shape)].view(*shape)
          for name, off, shape in param_layout(spec)}


def init_params(spec: PixelSpec, seed: int = 1, dtype=torch.float32):
  """Conv: Keras defaults (glorot_uniform kernel, zero bias); dense: 'normal'
  = N(0, 0.05) kernel and bias (dense_resnet_value.py:39-45)."""
  g = torch.Generator().manual_seed(seed)
  flat = torch.zeros(param_count(spec), dtype=torch.float64)
  for name, off, shape in param_layout(spec):
    n = math.prod(shape)
    if name.startswith('conv') and name.endswith('_k'):
      fan_in, fan_out = 9 * shape[2], 9 * shape[3]
      lim = math.sqrt(6.0 / (fan_in + fan_out))
      flat[off:off + n] = (torch.rand(n, generator=g, dtype=torch.float64)
                           * 2 - 1) * lim
    elif name.startswith('conv'):
      pass
    else:
      flat[off:off + n] = torch.randn(n, generator=g,
                                      dtype=torch.float64) * 0.05
  return flat.to(dtype)


def preprocess(images: torch.Tensor, spec: PixelSpec) -> torch.Tensor:
  """[B,T,H,W,C] u8 or float -> [B, th, tw, C*T] float (NHWC)."""
  if images.dtype == torch.uint8:
    images = images.to(torch.float32) * (1.0 / 255.0)
  b = images.shape[0]
  # image_prepro.py:24-28: a RAW reshape, no transpose (nw/nh names are
  # swapped in the reference but the reshape is memory-order preserving).
  x = images.reshape(b, spec.height, spec.width, spec.channels * spec.seq_len)
  x = x.permute(0, 3, 1, 2)
  x = F.interpolate(x, size=(spec.target_height, spec.target_width),
                    mode='bilinear', align_corners=False, antialias=False)
  return x.permute(0, 2, 3, 1).contiguous()


def encode(flat, spec: PixelSpec, images) -> torch.Tensor:
  """pixel_ebm.py:79-96 -> [B, 256]."""
  p = unpack(flat, spec)
  x = preprocess(images, spec).permute(0, 3, 1, 2)          # NCHW for torch
  for i in range(4):
    k = p[f'conv{i + 1}_k'].permute(3, 2, 0, 1)              # -> [co,ci,kh,kw]
    x = F.relu(F.conv2d(x, k, p[f'conv{i + 1}_b'], padding=1))
    x = F.max_pool2d(x, 2)                                   # 'valid'
  return x.mean(dim=(2, 3))


def value(flat, spec: PixelSpec, enc_tiled, act) -> torch.Tensor:
  """dense_resnet_value.py:31-36, 63-69 on x=[enc, act] -> E[R]."""
  p = unpack(flat, spec)
  x = torch.cat([enc_tiled, act], dim=-1)                    # pixel_ebm.py:117
  x = x @ p['v_dense0_W'] + p['v_dense0_b']
  for j in range(spec.value_blocks):
    y = torch.relu(x) @ p[f'v_blk{j}_d0_W'] + p[f'v_blk{j}_d0_b']
    y = torch.relu(y) @ p[f'v_blk{j}_d1_W'] + p[f'v_blk{j}_d1_b']
    y = torch.relu(y) @ p[f'v_blk{j}_d2_W'] + p[f'v_blk{j}_d2_b']
    x = x + y
  return x @ p['v_dense1_W'] + p['v_dense1_b']


def energy(flat, spec: PixelSpec, images, act) -> torch.Tensor:
  """PixelEBM.call without observation_encoding (pixel_ebm.py:98-125)."""
  enc = encode(flat, spec, images)
  n = act.shape[0] // images.shape[0]
  return value(flat, spec, torch.repeat_interleave(enc, n, dim=0), act)
