"""MLPEBM on the B200 engine -- same constructor and call protocol as the
reference `networks/mlp_ebm.py:25-96` (ResNetPreActivation body,
`networks/layers/resnet.py:32-153`; DenseSN weights,
`networks/layers/spectral_norm.py:31-86`).

The network owns one flat fp32 parameter tensor on the GPU (layout documented
in include/ibc_b200.h) and an `EBMEngine`.  `net((obs, act), training=...)`
returns `(energies[R], network_state)` like the reference.  Every number comes
from the CUDA library; without a GPU, construction raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin
from ibc_b200.ibc import specs


def _spec_dims(obs_spec) -> int:
  """T * sum of feature dims over the (dict of) observation specs."""
  total = 0
  for s in specs.flatten(obs_spec):
    n = 1
    for d in s.shape:
      n *= int(d)
    total += n
  return total


@gin.configurable('ResNetLayer')
def _resnet_layer_options(normalizer=None):
  """gin name `ResNetLayer.normalizer` (resnet.py:28-45).  Only None is on the
  hot path; Batch/LayerNorm are out of scope."""
  if normalizer is not None:
    raise NotImplementedError('ResNetLayer.normalizer=%r is not supported' % (normalizer,))
  return normalizer


@gin.configurable
class MLPEBM:
  """MLP-EBM; see module docstring."""

  def __init__(self,
               obs_spec,
               action_spec,
               width=512,
               depth=2,
               rate=0.1,
               name='MLPEBM',
               activation='relu',
               layers='MLPDropout',
               kernel_initializer='normal',
               bias_initializer='normal',
               dense_layer_type='regular',
               device=None,
               seed=None):
    if dense_layer_type not in ('regular', 'spectral_norm'):
      raise ValueError('Unexpected dense layer type')          # mlp_ebm.py:49
    if layers != 'ResNetPreActivation':
      raise NotImplementedError(
          "only layers='ResNetPreActivation' is built (every EBM gin selects it); got %r"
          % (layers,))
    if activation != 'relu':
      raise NotImplementedError("only activation='relu' is built; got %r" % (activation,))
    if rate != 0.0:
      raise NotImplementedError('dropout rate must be 0.0 (every EBM gin sets it); got %r' % rate)
    if kernel_initializer != 'normal' or bias_initializer != 'normal':
      raise NotImplementedError("only the 'normal' initializers are built")
    _resnet_layer_options()
    assert depth % 2 == 0, 'ResNet wants layers to be even numbers'   # resnet.py:42
    # obs_spec is (observation_spec, action_spec) as in get_cloning_network.py:75-78
    ob_spec, act_in_spec = obs_spec
    self.name = name
    self.state_spec = ()
    self.obs_dim = _spec_dims(ob_spec)
    self.act_dim = int(act_in_spec.shape[-1])
    if int(action_spec.shape[-1]) != 1:
      raise NotImplementedError('MLPEBM projects to a single energy (action_spec.shape[-1]==1)')
    self.width, self.depth = int(width), int(depth)
    self.dense_layer_type = dense_layer_type
    self.engine = eng.EBMEngine(self.obs_dim, self.act_dim, self.width, self.depth, device)
    self.device = self.engine.device
    g = torch.Generator(device='cpu')
    if seed is not None:
      g.manual_seed(int(seed))
    # Keras 'normal' = RandomNormal(0, 0.05) for kernels AND biases (mlp_ebm.py:38-39)
    flat = torch.randn(self.engine.param_count, generator=g, dtype=torch.float32) * 0.05
    self.params = flat.to(self.device).contiguous()
    self.built = True
    self.losses = []            # no regularisation losses (ibc_agent.py:249)
    self._sn_u: Optional[Dict[str, torch.Tensor]] = None
    self._effective = None
    if dense_layer_type == 'spectral_norm':
      self._sn_u = {}
      for nm, _, shape in self.param_layout():
        if nm.startswith('W') or nm == 'we':
          out = 1 if nm == 'we' else shape[-1]
          self._sn_u[nm] = torch.randn(1, out, generator=g, dtype=torch.float32).to(self.device)
    self._rebind()

  # -- parameters ---------------------------------------------------------------
  def param_layout(self):
    """[(name, offset, shape)] of the flat vector (include/ibc_b200.h)."""
    w, out, off = self.width, [], 0

    def add(nm, shape):
      nonlocal off
      out.append((nm, off, tuple(shape)))
      off += math.prod(shape)
    add('Wp', (self.obs_dim + self.act_dim, w))
    add('bp', (w,))
    for j in range(self.depth // 2):
      add('W1_%d' % j, (w, w)); add('b1_%d' % j, (w,))
      add('W2_%d' % j, (w, w)); add('b2_%d' % j, (w,))
    add('we', (w,))
    add('be', (1,))
    return out

  @property
  def variables(self):
    return [self.params]

  trainable_variables = variables

  def create_variables(self, *unused_args, **unused_kwargs):
    return None

  def set_params(self, flat: torch.Tensor):
    self.params.copy_(flat.to(self.device, torch.float32).reshape(-1))
    self.params_changed()

  def params_changed(self):
    """Call after mutating `self.params` in place (optimizer step)."""
    self._rebind()

  def _spectral_effective(self, update_u: bool, need_graph: bool):
    """W_bar = W / sigma with one power iteration (spectral_norm.py:59-80).
    O(W^2) per layer: done with torch ops on the GPU, kernels see W_bar."""
    raw = self.params.detach().clone().requires_grad_(need_graph)
    pieces = []
    for nm, off, shape in self.param_layout():
      t = raw[off:off + math.prod(shape)].view(*shape)
      if nm in self._sn_u:
        w2d = t.view(-1, 1) if nm == 'we' else t
        u = self._sn_u[nm]
        v = u @ w2d.t()
        v = v / (torch.sum(v ** 2) ** 0.5 + 1e-12)
        un = v @ w2d
        un = un / (torch.sum(un ** 2) ** 0.5 + 1e-12)
        sigma = (v @ w2d) @ un.t()
        t = (w2d / sigma).reshape(shape)
        # resnet.py:136 calls the projection layer WITHOUT `training=`, so DenseSN never
        # assigns its u (spectral_norm.py:76-80): the first layer keeps its initial u for the
        # whole run and is still normalised with it
        if update_u and nm != 'Wp':
          self._sn_u[nm] = un.detach()
      pieces.append(t.reshape(-1))
    return raw, torch.cat(pieces)

  def _set_effective(self, eff: torch.Tensor):
    """One persistent buffer for W_bar: the engine keys its captured Langevin graphs on the
    bound pointer, so a fresh tensor per optimizer step would defeat (and leak) them."""
    if self._effective is None or self._effective is self.params:
      self._effective = eff.detach().clone().contiguous()
      self.engine.bind_params(self._effective)
    else:
      self._effective.copy_(eff.detach())
      self.engine.params_changed()

  def _rebind(self):
    if self._sn_u is None:
      self._effective = self.params
      self.engine.bind_params(self._effective)
    else:
      _, eff = self._spectral_effective(update_u=False, need_graph=False)
      self._set_effective(eff)

  def effective_params_for_training(self):
    """(raw_leaf, effective_flat_with_graph) -- used by the agent to chain
    dL/dW_bar back to dL/dW through the power iteration (TF back-propagates
    through it; spectral_norm.py has no stop_gradient) and to update `u`
    (training=True only, spectral_norm.py:76-80)."""
    if self._sn_u is None:
      return None, self.params
    raw, eff = self._spectral_effective(update_u=True, need_graph=True)
    self._set_effective(eff)
    return raw, eff

  # -- forward ------------------------------------------------------------------
  def _flat_obs(self, obs) -> torch.Tensor:
    x = specs.flatten_observation(obs)
    return x.to(self.device, torch.float32).contiguous()

  def __call__(self, inputs, training=False, step_type=(), network_state=()):
    del training, step_type
    obs, act = inputs
    x = self._flat_obs(obs)
    act = act.to(self.device, torch.float32).contiguous()
    if x.shape[0] != act.shape[0]:
      raise ValueError('observations and actions must have the same leading dimension '
                       '(obs %d, act %d); tile observations first' % (x.shape[0], act.shape[0]))
    return self.engine.energy(x, act, 1), network_state

  call = __call__

  def energy_untiled(self, obs_flat: torch.Tensor, act: torch.Tensor, n: int) -> torch.Tensor:
    """E for obs [B, O] against act [B*n, A] without materialising the tiling."""
    return self.engine.energy(obs_flat, act, n)
