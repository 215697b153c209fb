"""Oracle (test infrastructure): MLPEBM with the ResNetPreActivation body.

Restates, in PyTorch-CPU:
  networks/mlp_ebm.py:72-96             MLPEBM.call
  networks/layers/resnet.py:135-153     ResNetPreActivationLayer.call
  networks/layers/spectral_norm.py:57-86 DenseSN.call (weight normalisation)
for the only configuration the EBM gins use: activation='relu',
normalizer=None, rate=0.0 (dropout is the identity).

Flat parameter layout (shared with the CUDA library, include/ibc_b200.h):
  Wp [d_in, W] row-major (Keras kernel = [in, out]), bp [W],
  then per residual block j: W1_j [W, W], b1_j [W], W2_j [W, W], b2_j [W],
  then we [W] (Keras kernel [W, 1]) and be [1].
d_in = obs_dim + act_dim with the observation columns FIRST (mlp_ebm.py:85).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Tuple

import torch


@dataclasses.dataclass(frozen=True)
class MLPSpec:
  obs_dim: int   # T * sum(obs key dims), already flattened (mlp_ebm.py:78-82)
  act_dim: int
  width: int
  depth: int     # number of W x W dense layers; depth/2 residual blocks

  def __post_init__(self):
    # resnet.py:42 -- "ResNet wants layers to be even numbers"
    assert self.depth % 2 == 0, 'depth must be even'

  @property
  def d_in(self) -> int:
    return self.obs_dim + self.act_dim

  @property
  def num_blocks(self) -> int:
    return self.depth // 2


def param_layout(spec: MLPSpec) -> List[Tuple[str, int, Tuple[int, ...]]]:
  """[(name, offset, shape)] of every tensor in the flat parameter vector."""
  out = []
  off = 0

  def add(name, shape):
    nonlocal off
    n = 1
    for s in shape:
      n *= s
    out.append((name, off, tuple(shape)))
    off += n

  w = spec.width
  add('Wp', (spec.d_in, w))
  add('bp', (w,))
  for j in range(spec.num_blocks):
    add(f'W1_{j}', (w, w))
    add(f'b1_{j}', (w,))
    add(f'W2_{j}', (w, w))
    add(f'b2_{j}', (w,))
  add('we', (w,))
  add('be', (1,))
  return out


def param_count(spec: MLPSpec) -> int:
  name, off, shape = param_layout(spec)[-1]
  return off + 1


def unpack(flat: torch.Tensor, spec: MLPSpec) -> Dict[str, torch.Tensor]:
  assert flat.numel() == param_count(spec), (flat.numel(), param_count(spec))
  d = {}
  for name, off, shape in param_layout(spec):
    n = 1
    for s in shape:
      n *= s
    d[name] = flat[off:off + n].view(*shape)
  return d


def init_params(spec: MLPSpec, seed: int = 1,
                dtype=torch.float32) -> torch.Tensor:
  """kernel_initializer='normal' and bias_initializer='normal' are Keras
  RandomNormal(mean=0, stddev=0.05) for kernels AND biases (mlp_ebm.py:38-39)."""
  g = torch.Generator().manual_seed(seed)
  return (torch.randn(param_count(spec), generator=g, dtype=torch.float64)
          * 0.05).to(dtype)


def energy(flat: torch.Tensor, spec: MLPSpec, obs: torch.Tensor,
           act: torch.Tensor) -> torch.Tensor:
  """E[R] for obs [R, O] (already tiled, flattened) and act [R, A]."""
  p = unpack(flat, spec)
  x = torch.cat([obs, act], dim=-1)               # mlp_ebm.py:85
  h = x @ p['Wp'] + p['bp']                       # resnet.py:136
  for j in range(spec.num_blocks):                # resnet.py:139-152
    v = torch.relu(h) @ p[f'W1_{j}'] + p[f'b1_{j}']
    h = h + torch.relu(v) @ p[f'W2_{j}'] + p[f'b2_{j}']
  e = h @ p['we'] + p['be']                       # mlp_ebm.py:91 (Dense(1))
  return e                                        # squeeze, mlp_ebm.py:94


def energy_and_dact_manual(flat: torch.Tensor, spec: MLPSpec,
                           obs: torch.Tensor, act: torch.Tensor):
  """(E[R], dE/dact[R, A]) by the hand-derived reverse chain the CUDA kernels
  implement: g = we; per block (reverse) g += m_h * ((m_v * (g W2^T)) W1^T);
  dE/dx = g Wp^T.  Cross-checked against autograd in tests."""
  p = unpack(flat, spec)
  x = torch.cat([obs, act], dim=-1)
  h = x @ p['Wp'] + p['bp']
  masks = []
  for j in range(spec.num_blocks):
    mh = (h > 0).to(h.dtype)
    v = torch.relu(h) @ p[f'W1_{j}'] + p[f'b1_{j}']
    mv = (v > 0).to(h.dtype)
    h = h + torch.relu(v) @ p[f'W2_{j}'] + p[f'b2_{j}']
    masks.append((mh, mv))
  e = h @ p['we'] + p['be']
  g = p['we'].unsqueeze(0).expand(x.shape[0], -1)
  for j in reversed(range(spec.num_blocks)):
    mh, mv = masks[j]
    u = (g @ p[f'W2_{j}'].t()) * mv
    g = g + (u @ p[f'W1_{j}'].t()) * mh
  dx = g @ p['Wp'].t()
  return e, dx[:, spec.obs_dim:]


# ---------------------------------------------------------------------------
# Spectral normalisation (D4RL config only).
# ---------------------------------------------------------------------------
def _l2normalize(v, eps=1e-12):
  # spectral_norm.py:59-60
  return v / (torch.sum(v ** 2) ** 0.5 + eps)


def spectral_normalize(w: torch.Tensor, u: torch.Tensor):
  """(W_bar, u_new) for kernel w [in, out] and u [1, out]
  (spectral_norm.py:61-74).  Differentiable w.r.t. w, like the reference
  (TF back-propagates through the power iteration)."""
  v = _l2normalize(u @ w.t())
  u_new = _l2normalize(v @ w)
  sigma = (v @ w) @ u_new.t()
  return w / sigma, u_new


def sn_kernel_names(spec: MLPSpec) -> List[str]:
  names = ['Wp']
  for j in range(spec.num_blocks):
    names += [f'W1_{j}', f'W2_{j}']
  names.append('we')
  return names


def init_sn_u(spec: MLPSpec, seed: int = 3, dtype=torch.float32):
  """u ~ N(0, 1), shape [1, out] per dense layer (spectral_norm.py:47-51)."""
  g = torch.Generator().manual_seed(seed)
  us = {}
  for name in sn_kernel_names(spec):
    out = 1 if name == 'we' else spec.width
    us[name] = torch.randn(1, out, generator=g, dtype=torch.float64).to(dtype)
  return us


def sn_effective_params(flat_raw: torch.Tensor, us, spec: MLPSpec):
  """Flat vector with every kernel replaced by W_bar; also the u's after a training call
  (only written back by the caller when training=True, spectral_norm.py:76-80).  The first
  layer is the exception: ResNetPreActivationLayer.call invokes `self._projection_layer(x)`
  WITHOUT `training=` (networks/layers/resnet.py:136), DenseSN.call then sees training=None
  and never assigns u (spectral_norm.py:76), so Wp keeps its initial u for the whole run
  while still being normalised with it: new_us['Wp'] is the u that was passed in.
  """
  p = unpack(flat_raw, spec)
  pieces = []
  new_us = {}
  for name, off, shape in param_layout(spec):
    t = p[name]
    if name in us:
      w2d = t.view(-1, 1) if name == 'we' else t
      wbar, un = spectral_normalize(w2d, us[name])
      new_us[name] = us[name].detach() if name == 'Wp' else un.detach()
      t = wbar.reshape(shape)
    pieces.append(t.reshape(-1))
  return torch.cat(pieces), new_us
