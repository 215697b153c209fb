"""Torch-facing wrapper of one `ibc_handle` (include/ibc_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every number is
produced by the CUDA library.  All tensors must be CUDA, contiguous, fp32
(fp64 for DFO uniforms).  Nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ibc_b200 import _lib
from ibc_b200._lib import LangevinCfg, LossCfg, MlpArch, PixelArch


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_f32(t: torch.Tensor, name: str):
  if not t.is_cuda:
    raise ValueError('%s must be a CUDA tensor (no CPU fallback)' % name)
  if t.dtype != torch.float32:
    raise ValueError('%s must be float32, got %s' % (name, t.dtype))
  if not t.is_contiguous():
    raise ValueError('%s must be contiguous' % name)


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
  """The current torch stream of the current device as a cudaStream_t (the raw getter skips
  the Stream object and device-index plumbing: 17 us -> ~1 us per call on the sampler path)."""
  if _raw_stream is not None:
    return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def langevin_cfg(num_iterations=25, sampler_stepsize_init=1e-1,
                 sampler_stepsize_decay=0.8, noise_scale=1.0, grad_clip=None,
                 delta_action_clip=0.1, use_polynomial_rate=True,
                 sampler_stepsize_final=1e-5, sampler_stepsize_power=2.0,
                 grad_norm_type='inf', apply_exp=False) -> LangevinCfg:
  if grad_norm_type not in _lib.NORM_BY_NAME:
    raise KeyError(grad_norm_type)   # same failure mode as mcmc.py:197-200
  return LangevinCfg(int(num_iterations), float(sampler_stepsize_init),
                     float(sampler_stepsize_decay), float(noise_scale),
                     -1.0 if grad_clip is None else float(grad_clip),
                     float(delta_action_clip), int(bool(use_polynomial_rate)),
                     float(sampler_stepsize_final),
                     float(sampler_stepsize_power),
                     _lib.NORM_BY_NAME[grad_norm_type], int(bool(apply_exp)))


def loss_cfg(softmax_temperature=1.0, add_grad_penalty=True,
             grad_norm_type='inf', grad_margin=1.0, square_grad_penalty=True,
             grad_loss_weight=1.0, ebm_loss_type='info_nce') -> LossCfg:
  if ebm_loss_type not in _lib.LOSS_BY_NAME:
    raise ValueError('Unsupported EBM loss type.')          # ibc_agent.py:335
  return LossCfg(float(softmax_temperature), int(bool(add_grad_penalty)),
                 _lib.NORM_BY_NAME[grad_norm_type],
                 -1.0 if grad_margin is None else float(grad_margin),
                 int(bool(square_grad_penalty)), float(grad_loss_weight),
                 _lib.LOSS_BY_NAME[ebm_loss_type])


def param_count(obs_dim: int, act_dim: int, width: int, depth: int) -> int:
  """Pure arithmetic; usable without a GPU."""
  d_in = obs_dim + act_dim
  return d_in * width + width + (depth // 2) * (2 * width * width + 2 * width) + width + 1


class EBMEngine:
  """Owns an ibc_handle for one MLPEBM architecture on one GPU."""

  def __init__(self, obs_dim: int, act_dim: int, width: int, depth: int,
               device: Optional[torch.device] = None):
    if depth % 2 != 0:
      raise AssertionError('ResNet wants layers to be even numbers')  # resnet.py:42
    if not torch.cuda.is_available():
      raise _lib.IbcError('ibc_b200 needs a CUDA device (B200); there is no CPU path')
    self.lib = _lib.load()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
        else torch.device(device)
    self.arch = MlpArch(obs_dim, act_dim, width, depth)
    self.obs_dim, self.act_dim, self.width, self.depth = obs_dim, act_dim, width, depth
    self.param_count = int(self.lib.ibc_param_count(ctypes.byref(self.arch)))
    h = ctypes.c_void_p()
    _lib.check(self.lib.ibc_create(ctypes.byref(self.arch), self.device.index or 0,
                                   ctypes.byref(h)))
    self._h = h
    self._params = None

  def __del__(self):
    h = getattr(self, '_h', None)
    if h is not None and h.value:
      try:
        self.lib.ibc_destroy(h)
      except Exception:  # interpreter shutdown
        pass
      self._h = None

  # -- configuration ---------------------------------------------------------
  @property
  def precision(self) -> int:
    return int(self.lib.ibc_get_precision(self._h))

  def set_precision(self, precision: int):
    _lib.check(self.lib.ibc_set_precision(self._h, int(precision)), self._h)

  def bind_params(self, params: torch.Tensor):
    _check_f32(params, 'params')
    if params.numel() != self.param_count:
      raise ValueError('params has %d elements, expected %d' % (params.numel(), self.param_count))
    self._params = params
    _lib.check(self.lib.ibc_bind_params(self._h, _ptr(params), _stream()), self._h)

  def params_changed(self):
    _lib.check(self.lib.ibc_params_changed(self._h), self._h)

  def kernel_status(self) -> int:
    st = ctypes.c_int32(0)
    _lib.check(self.lib.ibc_kernel_status(self._h, ctypes.byref(st)), self._h)
    return int(st.value)

  def set_profiling(self, on: bool):
    _lib.check(self.lib.ibc_set_profiling(self._h, int(bool(on))), self._h)

  def train_timings(self):
    """ms of {forward, info_nce, reverse chain, weight gradient, first/last layer}."""
    out = (ctypes.c_float * 5)()
    _lib.check(self.lib.ibc_get_train_timings(self._h, out), self._h)
    return [float(x) for x in out]

  # -- energy ------------------------------------------------------------------
  def _check_obs_act(self, obs, act, n):
    _check_f32(obs, 'obs')
    _check_f32(act, 'act')
    if obs.dim() != 2 or obs.shape[1] != self.obs_dim:
      raise ValueError('obs must be [B, %d], got %s' % (self.obs_dim, tuple(obs.shape)))
    b = obs.shape[0]
    if act.dim() != 2 or act.shape[1] != self.act_dim or act.shape[0] != b * n:
      raise ValueError('act must be [%d, %d], got %s' % (b * n, self.act_dim, tuple(act.shape)))
    return b

  def energy(self, obs: torch.Tensor, act: torch.Tensor, n: int) -> torch.Tensor:
    b = self._check_obs_act(obs, act, n)
    e = torch.empty(b * n, device=obs.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_mlp_energy(self._h, _ptr(obs), b, _ptr(act), n, _ptr(e), _stream()),
               self._h)
    return e

  def energy_dact(self, obs, act, n):
    b = self._check_obs_act(obs, act, n)
    e = torch.empty(b * n, device=obs.device, dtype=torch.float32)
    g = torch.empty(b * n, self.act_dim, device=obs.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_mlp_energy_dact(self._h, _ptr(obs), b, _ptr(act), n, _ptr(e),
                                            _ptr(g), _stream()), self._h)
    return e, g

  # -- samplers ------------------------------------------------------------------
  def dfo(self, obs, actions, n, min_actions, max_actions, temperature=1.0,
          num_iterations=3, iteration_std=0.33, uniforms=None, normals=None, seed=0):
    """In-place on `actions`; returns probs [B*n]."""
    b = self._check_obs_act(obs, actions, n)
    _check_f32(min_actions, 'min_actions')
    _check_f32(max_actions, 'max_actions')
    if uniforms is not None:
      if uniforms.dtype != torch.float64 or not uniforms.is_cuda or not uniforms.is_contiguous():
        raise ValueError('uniforms must be a contiguous CUDA float64 tensor')
      if uniforms.numel() != num_iterations * b * n:
        raise ValueError('uniforms must have num_iterations*B*n elements')
    if normals is not None:
      _check_f32(normals, 'normals')
      if normals.numel() != (num_iterations - 1) * b * n * self.act_dim:
        raise ValueError('normals must have (num_iterations-1)*B*n*A elements')
    probs = torch.empty(b * n, device=obs.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_dfo(self._h, _ptr(obs), b, _ptr(actions), n, _ptr(min_actions),
                                _ptr(max_actions), float(temperature), int(num_iterations),
                                float(iteration_std), _ptr(uniforms), _ptr(normals), int(seed),
                                _ptr(probs), _stream()), self._h)
    return probs

  def langevin(self, obs, actions, n, min_actions, max_actions, cfg: LangevinCfg,
               normals=None, seed=0, return_chain=False):
    """In-place on `actions`; returns (chain_actions, energies, grad_norms) or None."""
    b = self._check_obs_act(obs, actions, n)
    _check_f32(min_actions, 'min_actions')
    _check_f32(max_actions, 'max_actions')
    r, k = b * n, cfg.num_iterations
    if normals is not None:
      _check_f32(normals, 'normals')
      if normals.numel() != k * r * self.act_dim:
        raise ValueError('normals must have K*B*n*A elements')
    ca = ce = cg = None
    if return_chain:
      ca = torch.empty(k, r, self.act_dim, device=obs.device, dtype=torch.float32)
      ce = torch.empty(k, r, device=obs.device, dtype=torch.float32)
      cg = torch.empty(k, r, device=obs.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_langevin(self._h, _ptr(obs), b, _ptr(actions), n, _ptr(min_actions),
                                     _ptr(max_actions), ctypes.byref(cfg), _ptr(normals),
                                     int(seed), _ptr(ca), _ptr(ce), _ptr(cg), _stream()),
               self._h)
    return (ca, ce, cg) if return_chain else None

  # -- training --------------------------------------------------------------------
  def train_grads(self, obs, actions, num_counter_examples, cfg: LossCfg, global_batch,
                  grads: Optional[torch.Tensor] = None, want_energies=False):
    """Returns (per_example_loss [B], grad_loss [B], energies or None, grads)."""
    n1 = num_counter_examples + 1
    b = self._check_obs_act(obs, actions, n1)
    if grads is None:
      grads = torch.empty(self.param_count, device=obs.device, dtype=torch.float32)
    _check_f32(grads, 'grads')
    per_example = torch.empty(b, device=obs.device, dtype=torch.float32)
    grad_loss = torch.empty(b, device=obs.device, dtype=torch.float32)
    energies = torch.empty(b * n1, device=obs.device, dtype=torch.float32) if want_energies else None
    _lib.check(self.lib.ibc_train_grads(self._h, _ptr(obs), b, _ptr(actions),
                                        int(num_counter_examples), ctypes.byref(cfg),
                                        int(global_batch), _ptr(per_example), _ptr(grad_loss),
                                        _ptr(energies), _ptr(grads), _stream()), self._h)
    return per_example, grad_loss, energies, grads

  def stream_wait_loss(self, stream: torch.cuda.Stream) -> bool:
    """If the last train_grads recorded its loss event (per_example_loss final, backward still
    to run), `stream` waits for it and True is returned."""
    rec = ctypes.c_int32(0)
    _lib.check(self.lib.ibc_stream_wait_loss(self._h, ctypes.c_void_p(stream.cuda_stream),
                                             ctypes.byref(rec)), self._h)
    return bool(rec.value)


class PixelEngine:
  """Owns an ibc_pixel_handle (late-fusion PixelEBM: ConvMaxpool encoder +
  DenseResnetValue head) on one GPU."""

  def __init__(self, seq_len, height, width, channels, target_height, target_width, act_dim,
               value_width, value_blocks, device: Optional[torch.device] = None):
    if not torch.cuda.is_available():
      raise _lib.IbcError('ibc_b200 needs a CUDA device (B200); there is no CPU path')
    self.lib = _lib.load()
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None \
        else torch.device(device)
    self.arch = PixelArch(seq_len, height, width, channels, target_height, target_width, act_dim,
                          value_width, value_blocks)
    self.act_dim = act_dim
    self.param_count = int(self.lib.ibc_pixel_param_count(ctypes.byref(self.arch)))
    h = ctypes.c_void_p()
    _lib.check(self.lib.ibc_pixel_create(ctypes.byref(self.arch), self.device.index or 0,
                                         ctypes.byref(h)), None, pixel=True)
    self._h = h
    self._params = None

  def __del__(self):
    h = getattr(self, '_h', None)
    if h is not None and h.value:
      try:
        self.lib.ibc_pixel_destroy(h)
      except Exception:
        pass
      self._h = None

  def bind_params(self, params: torch.Tensor):
    _check_f32(params, 'params')
    if params.numel() != self.param_count:
      raise ValueError('params has %d elements, expected %d' % (params.numel(), self.param_count))
    self._params = params
    _lib.check(self.lib.ibc_pixel_bind_params(self._h, _ptr(params)), self._h, pixel=True)

  @property
  def precision(self) -> int:
    return int(self.lib.ibc_pixel_get_precision(self._h))

  def set_precision(self, precision: int):
    _lib.check(self.lib.ibc_pixel_set_precision(self._h, int(precision)), self._h, pixel=True)

  def kernel_status(self) -> int:
    st = ctypes.c_int32(-1)
    _lib.check(self.lib.ibc_pixel_kernel_status(self._h, ctypes.byref(st)), self._h, pixel=True)
    return int(st.value)

  def _check_images(self, images):
    a = self.arch
    if not images.is_cuda or not images.is_contiguous():
      raise ValueError('images must be a contiguous CUDA tensor')
    if images.dtype not in (torch.uint8, torch.float32):
      raise ValueError('images must be uint8 or float32')
    if tuple(images.shape[1:]) != (a.seq_len, a.height, a.width, a.channels):
      raise ValueError('images must be [B, %d, %d, %d, %d], got %s'
                       % (a.seq_len, a.height, a.width, a.channels, tuple(images.shape)))
    return images.shape[0], int(images.dtype == torch.uint8)

  def encode(self, images: torch.Tensor) -> torch.Tensor:
    b, is_u8 = self._check_images(images)
    enc = torch.empty(b, 256, device=images.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_encode(self._h, _ptr(images), is_u8, b, _ptr(enc), _stream()),
               self._h, pixel=True)
    return enc

  def energy(self, enc: torch.Tensor, act: torch.Tensor, n: int) -> torch.Tensor:
    _check_f32(enc, 'enc')
    _check_f32(act, 'act')
    b = enc.shape[0]
    if enc.shape[1] != 256 or tuple(act.shape) != (b * n, self.act_dim):
      raise ValueError('enc must be [B, 256] and act [B*n, A]')
    e = torch.empty(b * n, device=enc.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_energy(self._h, _ptr(enc), b, _ptr(act), n, _ptr(e), _stream()),
               self._h, pixel=True)
    return e

  def dfo(self, enc, actions, n, min_actions, max_actions, temperature=1.0, num_iterations=3,
          iteration_std=0.33, uniforms=None, normals=None, seed=0):
    _check_f32(enc, 'enc')
    _check_f32(actions, 'actions')
    b = enc.shape[0]
    if tuple(actions.shape) != (b * n, self.act_dim):
      raise ValueError('actions must be [B*n, A]')
    if uniforms is not None and (uniforms.dtype != torch.float64 or not uniforms.is_cuda
                                 or uniforms.numel() != num_iterations * b * n):
      raise ValueError('uniforms must be CUDA float64 with num_iterations*B*n elements')
    if normals is not None:
      _check_f32(normals, 'normals')
    probs = torch.empty(b * n, device=enc.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_dfo(self._h, _ptr(enc), b, _ptr(actions), n, _ptr(min_actions),
                                      _ptr(max_actions), float(temperature), int(num_iterations),
                                      float(iteration_std), _ptr(uniforms), _ptr(normals),
                                      int(seed), _ptr(probs), _stream()), self._h, pixel=True)
    return probs

  def energy_dact(self, enc: torch.Tensor, act: torch.Tensor, n: int):
    """(energies [B*n], +dE/dact [B*n, A]) of the value head on an un-tiled encoding."""
    _check_f32(enc, 'enc')
    _check_f32(act, 'act')
    b = enc.shape[0]
    if enc.shape[1] != 256 or tuple(act.shape) != (b * n, self.act_dim):
      raise ValueError('enc must be [B, 256] and act [B*n, A]')
    e = torch.empty(b * n, device=enc.device, dtype=torch.float32)
    g = torch.empty(b * n, self.act_dim, device=enc.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_energy_dact(self._h, _ptr(enc), b, _ptr(act), n, _ptr(e),
                                              _ptr(g), _stream()), self._h, pixel=True)
    return e, g

  def langevin(self, enc, actions, n, min_actions, max_actions, cfg: LangevinCfg,
               normals=None, seed=0, return_chain=False):
    """Late-fusion Langevin chain, in place on `actions`; returns
    (chain_actions, energies, grad_norms) or None."""
    _check_f32(enc, 'enc')
    _check_f32(actions, 'actions')
    _check_f32(min_actions, 'min_actions')
    _check_f32(max_actions, 'max_actions')
    b = enc.shape[0]
    if enc.shape[1] != 256 or tuple(actions.shape) != (b * n, self.act_dim):
      raise ValueError('enc must be [B, 256] and actions [B*n, A]')
    r, k = b * n, cfg.num_iterations
    if normals is not None:
      _check_f32(normals, 'normals')
      if normals.numel() != k * r * self.act_dim:
        raise ValueError('normals must have K*B*n*A elements')
    ca = ce = cg = None
    if return_chain:
      ca = torch.empty(k, r, self.act_dim, device=enc.device, dtype=torch.float32)
      ce = torch.empty(k, r, device=enc.device, dtype=torch.float32)
      cg = torch.empty(k, r, device=enc.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_langevin(self._h, _ptr(enc), b, _ptr(actions), n,
                                           _ptr(min_actions), _ptr(max_actions),
                                           ctypes.byref(cfg), _ptr(normals), int(seed),
                                           _ptr(ca), _ptr(ce), _ptr(cg), _stream()),
               self._h, pixel=True)
    return (ca, ce, cg) if return_chain else None

  def train_begin(self, images, num_counter_examples) -> torch.Tensor:
    """First half of train_grads: the encoder pass (activations stay in the handle);
    returns the encoding [B, 256] for the counter-example sampler."""
    b, is_u8 = self._check_images(images)
    enc = torch.empty(b, 256, device=images.device, dtype=torch.float32)
    _lib.check(self.lib.ibc_pixel_train_begin(self._h, _ptr(images), is_u8, b,
                                              int(num_counter_examples), _ptr(enc), _stream()),
               self._h, pixel=True)
    self._open = (b, int(num_counter_examples))
    return enc

  def train_finish(self, actions, cfg: LossCfg, global_batch,
                   grads: Optional[torch.Tensor] = None, want_energies=False):
    """Second half: (per_example_loss, grad_loss, energies or None, grads)."""
    opened = getattr(self, '_open', None)
    if opened is None:
      raise _lib.IbcError('train_finish without train_begin')
    self._open = None
    b, n = opened
    _check_f32(actions, 'actions')
    if tuple(actions.shape) != (b * (n + 1), self.act_dim):
      raise ValueError('actions must be [B*(N+1), A]')
    if grads is None:
      grads = torch.empty(self.param_count, device=actions.device, dtype=torch.float32)
    per_example = torch.empty(b, device=actions.device, dtype=torch.float32)
    grad_loss = torch.empty(b, device=actions.device, dtype=torch.float32)
    energies = torch.empty(b * (n + 1), device=actions.device, dtype=torch.float32) \
        if want_energies else None
    _lib.check(self.lib.ibc_pixel_train_finish(self._h, _ptr(actions), ctypes.byref(cfg),
                                               int(global_batch), _ptr(per_example),
                                               _ptr(grad_loss), _ptr(energies), _ptr(grads),
                                               _stream()), self._h, pixel=True)
    return per_example, grad_loss, energies, grads

  def train_grads(self, images, actions, num_counter_examples, cfg: LossCfg, global_batch,
                  grads: Optional[torch.Tensor] = None, want_energies=False,
                  want_grad_loss=False):
    """Returns (per_example_loss [B], energies or None, grads), or with
    want_grad_loss (per_example_loss, grad_loss [B], energies or None, grads)."""
    b, is_u8 = self._check_images(images)
    n1 = num_counter_examples + 1
    _check_f32(actions, 'actions')
    if tuple(actions.shape) != (b * n1, self.act_dim):
      raise ValueError('actions must be [B*(N+1), A]')
    if grads is None:
      grads = torch.empty(self.param_count, device=images.device, dtype=torch.float32)
    per_example = torch.empty(b, device=images.device, dtype=torch.float32)
    energies = torch.empty(b * n1, device=images.device, dtype=torch.float32) if want_energies else None
    grad_loss = torch.empty(b, device=images.device, dtype=torch.float32) if want_grad_loss else None
    _lib.check(self.lib.ibc_pixel_train_grads(self._h, _ptr(images), is_u8, b, _ptr(actions),
                                              int(num_counter_examples), ctypes.byref(cfg),
                                              int(global_batch), _ptr(per_example),
                                              _ptr(grad_loss), _ptr(energies), _ptr(grads),
                                              _stream()),
               self._h, pixel=True)
    if want_grad_loss:
      return per_example, grad_loss, energies, grads
    return per_example, energies, grads


# -- handle-free ops -----------------------------------------------------------------
def resample(logits: torch.Tensor, uniforms: torch.Tensor, want_draws=False, want_counts=False):
  """categorical_bincount + tf.repeat (mcmc.py:32-55,136-137).  Returns
  (rows [B*count] int64, draws or None, counts or None)."""
  _check_f32(logits, 'logits')
  if uniforms.dtype != torch.float64 or not uniforms.is_cuda or not uniforms.is_contiguous():
    raise ValueError('uniforms must be a contiguous CUDA float64 tensor')
  b, n = logits.shape
  count = uniforms.shape[1]
  if uniforms.shape[0] != b:
    raise ValueError('uniforms must be [B, count]')
  rows = torch.empty(b * count, device=logits.device, dtype=torch.int64)
  draws = torch.empty(b, count, device=logits.device, dtype=torch.int32) if want_draws else None
  counts = torch.empty(b, n, device=logits.device, dtype=torch.int32) if want_counts else None
  _lib.check(_lib.load().ibc_resample(_ptr(logits), _ptr(uniforms), b, n, count, _ptr(draws),
                                      _ptr(counts), _ptr(rows), _stream()))
  return rows, draws, counts


def softmax_rows(logits: torch.Tensor, temperature: float = 1.0) -> torch.Tensor:
  _check_f32(logits, 'logits')
  b, n = logits.shape
  probs = torch.empty_like(logits)
  _lib.check(_lib.load().ibc_softmax_rows(_ptr(logits), b, n, float(temperature), _ptr(probs),
                                          _stream()))
  return probs


def uniform_actions(rows: int, min_actions, max_actions, u=None, seed=0) -> torch.Tensor:
  _check_f32(min_actions, 'min_actions')
  _check_f32(max_actions, 'max_actions')
  a = min_actions.numel()
  if u is not None:
    _check_f32(u, 'u')
    if u.numel() != rows * a:
      raise ValueError('u must have rows*A elements')
  out = torch.empty(rows, a, device=min_actions.device, dtype=torch.float32)
  _lib.check(_lib.load().ibc_uniform_actions(_ptr(u), rows, a, _ptr(min_actions),
                                             _ptr(max_actions), int(seed), _ptr(out), _stream()))
  return out


def greedy_actions(probs, values, min_actions=None, max_actions=None):
  """probs [B, n], values [B, n, A] -> [B, A]: the value at the first arg max of every row,
  clipped to the bounds when given."""
  _check_f32(probs, 'probs')
  _check_f32(values, 'values')
  b, n = probs.shape
  a = values.shape[-1]
  if values.numel() != b * n * a:
    raise ValueError('values must be [B, n, A]')
  if min_actions is not None:
    _check_f32(min_actions, 'min_actions')
    _check_f32(max_actions, 'max_actions')
  out = torch.empty(b, a, device=probs.device, dtype=torch.float32)
  _lib.check(_lib.load().ibc_greedy_actions(_ptr(probs), _ptr(values), b, n, a, _ptr(min_actions),
                                            _ptr(max_actions), _ptr(out), _stream()))
  return out


def counter_examples_uniform(actions, n: int, min_actions, max_actions, u=None, seed=0):
  """[B * (n + 1), A]: n uniform negatives per row of `actions` (the samples of
  uniform_actions(B * n, ...) with the same u / seed) followed by the row itself."""
  _check_f32(actions, 'actions')
  _check_f32(min_actions, 'min_actions')
  _check_f32(max_actions, 'max_actions')
  b, a = actions.shape
  if u is not None:
    _check_f32(u, 'u')
    if u.numel() != b * n * a:
      raise ValueError('u must have B*n*A elements')
  out = torch.empty(b * (n + 1), a, device=actions.device, dtype=torch.float32)
  _lib.check(_lib.load().ibc_counter_examples_uniform(_ptr(u), b, n, a, _ptr(min_actions),
                                                      _ptr(max_actions), int(seed), _ptr(actions),
                                                      _ptr(out), _stream()))
  return out


def gather_perturb(src, rows, std, min_actions, max_actions, normals=None, seed=0):
  _check_f32(src, 'src')
  r, a = rows.numel(), src.shape[1]
  if rows.dtype != torch.int64 or not rows.is_cuda:
    raise ValueError('rows must be CUDA int64')
  if normals is not None:
    _check_f32(normals, 'normals')
  out = torch.empty(r, a, device=src.device, dtype=torch.float32)
  _lib.check(_lib.load().ibc_dfo_gather_perturb(_ptr(src), _ptr(rows), r, a, _ptr(normals),
                                                int(seed), float(std), _ptr(min_actions),
                                                _ptr(max_actions), _ptr(out), _stream()))
  return out


def langevin_update(actions, de_dact_pos, stepsize, min_actions, max_actions, noise_scale=1.0,
                    grad_clip=None, delta_action_clip=0.1, grad_norm_type='inf', normals=None,
                    seed=0):
  """One langevin_step update in place given +dE/dact; returns grad_norms [R]."""
  _check_f32(actions, 'actions')
  _check_f32(de_dact_pos, 'dE_dact')
  r, a = actions.shape
  if normals is not None:
    _check_f32(normals, 'normals')
  norms = torch.empty(r, device=actions.device, dtype=torch.float32)
  _lib.check(_lib.load().ibc_langevin_update(
      _ptr(actions), _ptr(de_dact_pos), r, a, _ptr(normals), int(seed), float(noise_scale),
      -1.0 if grad_clip is None else float(grad_clip), float(delta_action_clip), float(stepsize),
      _ptr(min_actions), _ptr(max_actions), _lib.NORM_BY_NAME[grad_norm_type], _ptr(norms),
      _stream()))
  return norms


def info_nce(predictions: torch.Tensor, temperature=1.0, grad_scale=1.0, want_grad=False):
  _check_f32(predictions, 'predictions')
  b, n1 = predictions.shape
  loss = torch.empty(b, device=predictions.device, dtype=torch.float32)
  dpred = torch.empty_like(predictions) if want_grad else None
  _lib.check(_lib.load().ibc_info_nce(_ptr(predictions), b, n1, float(temperature),
                                      float(grad_scale), _ptr(loss), _ptr(dpred), _stream()))
  return (loss, dpred) if want_grad else loss


def cd_loss(predictions: torch.Tensor, combined_actions: Optional[torch.Tensor], loss_type='cd',
            grad_scale=1.0, want_grad=False):
  """ebm_loss.cd / clipped_cd (ibc/losses/ebm_loss.py:51-66, 109-148) on predictions [B, n+1];
  combined_actions [B*(n+1), A] (counter examples, then the true action, per observation)."""
  _check_f32(predictions, 'predictions')
  b, n1 = predictions.shape
  a = 0
  if combined_actions is not None:
    _check_f32(combined_actions, 'combined_actions')
    if combined_actions.shape[0] != b * n1:
      raise ValueError('combined_actions must be [B * (n + 1), A]')
    a = combined_actions.shape[1]
  loss = torch.empty(b, device=predictions.device, dtype=torch.float32)
  dpred = torch.empty_like(predictions) if want_grad else None
  _lib.check(_lib.load().ibc_cd_loss(_ptr(predictions), _ptr(combined_actions), b, n1, a,
                                     _lib.LOSS_BY_NAME[loss_type], float(grad_scale), _ptr(loss),
                                     _ptr(dpred), _stream()))
  return (loss, dpred) if want_grad else loss


def grad_penalty_from_dact(de_dact, batch_size, cfg: LossCfg, grad_scale=1.0, want_cotangent=False):
  _check_f32(de_dact, 'dE_dact')
  r, a = de_dact.shape
  n1 = r // batch_size
  gl = torch.empty(batch_size, device=de_dact.device, dtype=torch.float32)
  cot = torch.empty_like(de_dact) if want_cotangent else None
  _lib.check(_lib.load().ibc_grad_penalty_from_dact(_ptr(de_dact), batch_size, n1, a,
                                                    ctypes.byref(cfg), float(grad_scale),
                                                    _ptr(gl), _ptr(cot), _stream()))
  return (gl, cot) if want_cotangent else gl


def adam_step(params, grads, m, v, lr_t, beta1=0.9, beta2=0.999, epsilon=1e-7):
  for t, name in ((params, 'params'), (grads, 'grads'), (m, 'm'), (v, 'v')):
    _check_f32(t, name)
  _lib.check(_lib.load().ibc_adam_step(_ptr(params), _ptr(grads), _ptr(m), _ptr(v),
                                       params.numel(), float(lr_t), float(beta1), float(beta2),
                                       float(epsilon), _stream()))


def langevin_stepsizes(cfg: LangevinCfg):
  out = (ctypes.c_float * max(cfg.num_iterations, 1))()
  _lib.check(_lib.load().ibc_langevin_stepsizes(ctypes.byref(cfg), out))
  return [float(out[i]) for i in range(cfg.num_iterations)]


def gemm_tf32(a: torch.Tensor, b: torch.Tensor, m: int, n: int, k: int, trans_a=False,
              trans_b=False, bias=None, gate_c=None, res=None, relu_a=False, relu_out=False,
              out: Optional[torch.Tensor] = None, check: bool = True):
  """C[m, n] = f(op(A) . op(B) + bias) (.) (gate_c > 0) + res) on the general tcgen05 tf32
  GEMM (ibc_gemm_tf32); A / B are 2-D row-major (their row stride is the leading
  dimension); with `out` the product is accumulated into it (split-K atomics).
  Returns (C, status); with check=False the call does not wait and status is None."""
  for t, name in ((a, 'A'), (b, 'B')):
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1):
      raise ValueError('%s must be a 2-D CUDA float32 tensor with unit column stride' % name)
  for t in (bias, gate_c, res):
    if t is not None:
      _check_f32(t, 'epilogue operand')
  c = out if out is not None else torch.empty(m, n, device=a.device, dtype=torch.float32)
  st = ctypes.c_int32(-1)
  _lib.check(_lib.load().ibc_gemm_tf32(
      _ptr(a), a.stride(0), int(trans_a), _ptr(b), b.stride(0), int(trans_b), _ptr(c), c.stride(0),
      m, n, k, _ptr(bias) if bias is not None else None,
      _ptr(gate_c) if gate_c is not None else None, _ptr(res) if res is not None else None,
      int(relu_a), int(relu_out), int(out is not None), ctypes.byref(st) if check else None,
      _stream()))
  return c, (int(st.value) if check else None)


def conv3x3_tf32(x, kernel, bias=None, relu=False, mode=0, dy=None, out=None):
  """The encoder's implicit-GEMM convolution on its own (ibc_conv3x3_tf32): mode 0 forward
  (x [nb,H,W,C], kernel [3,3,C,Co]) -> [nb,H,W,Co]; mode 1 weight gradient accumulated into
  `out` [3,3,C,Co] from x and dy [nb,H,W,Co]; mode 2 data gradient of dy -> [nb,H,W,C].
  Returns (result, status)."""
  _check_f32(kernel, 'kernel')
  c, co = int(kernel.shape[2]), int(kernel.shape[3])
  ref = x if x is not None else dy
  nb, h, w = int(ref.shape[0]), int(ref.shape[1]), int(ref.shape[2])
  for t in (x, dy, bias):
    if t is not None:
      _check_f32(t, 'operand')
  if out is None:
    shape = {0: (nb, h, w, co), 1: (3, 3, c, co), 2: (nb, h, w, c)}[mode]
    out = torch.zeros(shape, device=kernel.device, dtype=torch.float32)
  st = ctypes.c_int32(-1)
  _lib.check(_lib.load().ibc_conv3x3_tf32(_ptr(x), nb, h, w, c, _ptr(kernel), co, _ptr(bias),
                                          int(relu), int(mode), _ptr(dy), _ptr(out),
                                          ctypes.byref(st), _stream()))
  return out, int(st.value)


def umma_f16_selftest(a: torch.Tensor, b: torch.Tensor, lbo_a=0, lbo_b=0, sbo=0, layout_type=0,
                      kstep_bytes=0):
  """D[128,N] = A[K,128]^T . B[K,N] through tcgen05 kind::f16 with MN-major operand images;
  descriptor fields 0 = the library's own.  Returns (D, status)."""
  _check_f32(a, 'A')
  _check_f32(b, 'B')
  k, n = b.shape
  assert a.shape == (k, 128)
  d = torch.zeros(128, n, device=a.device, dtype=torch.float32)
  st = ctypes.c_int32(-1)
  _lib.check(_lib.load().ibc_umma_f16_selftest(_ptr(a), _ptr(b), n, k, lbo_a, lbo_b, sbo,
                                               layout_type, kstep_bytes, _ptr(d),
                                               ctypes.byref(st), _stream()))
  return d, int(st.value)


def umma_selftest(a: torch.Tensor, bt: torch.Tensor, mode: int = 0):
  """D = A[128,K] . Bt[N,K]^T through tcgen05; returns (D, status)."""
  _check_f32(a, 'A')
  _check_f32(bt, 'Bt')
  n, k = bt.shape
  d = torch.zeros(128, n, device=a.device, dtype=torch.float32)
  st = ctypes.c_int32(-1)
  _lib.check(_lib.load().ibc_umma_selftest(_ptr(a), _ptr(bt), n, k, int(mode), _ptr(d),
                                           ctypes.byref(st), _stream()))
  return d, int(st.value)
