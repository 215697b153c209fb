"""Late-fusion PixelEBM on the B200 engine -- same constructor and call
protocol as the reference `networks/pixel_ebm.py:46-125`
(ConvMaxpoolEncoder `networks/layers/conv_maxpool.py:20-48`, DenseResnetValue
`networks/layers/dense_resnet_value.py:22-69`, preprocessing
`networks/utils/image_prepro.py:20-43`).

  net.encode(obs, training)                      -> [B, 256]
  net((obs, act), training, observation_encoding=enc_tiled) -> (E [R], state)
  net((obs, act), training)                      -> encodes, tiles R//B, evaluates

Observations are dicts with an 'rgb' entry [B, T, H, W, 3] (uint8, or float32 in
[0, 1]); other keys are ignored like in the reference.
"""
from __future__ import annotations

import math

import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin

_CONV = (32, 64, 128, 256)


@gin.configurable
class DenseResnetValue:
  """Holds the gin-configurable hyper-parameters of the value head
  (dense_resnet_value.py:21-29)."""

  def __init__(self, width=512, num_blocks=2):
    self.width, self.num_blocks = int(width), int(num_blocks)


def get_encoder_network(encoder_network, target_height, target_width, channels):
  del target_height, target_width, channels
  if encoder_network == 'ConvMaxpoolEncoder':
    return 'ConvMaxpoolEncoder'
  if encoder_network == 'SpatialAttentionEncoder':
    raise NotImplementedError('SpatialAttentionEncoder is out of scope (no shipped gin selects it)')
  raise ValueError('Unsupported encoder_network %s' % encoder_network)      # pixel_ebm.py:36


def get_value_network(value_network):
  if value_network == 'DenseResnetValue':
    return DenseResnetValue()
  raise ValueError('Unsupported value_network %s' % value_network)           # pixel_ebm.py:43


@gin.configurable
class PixelEBM:
  """Late fusion PixelEBM."""

  def __init__(self, obs_spec, action_spec, encoder_network, value_network, target_height=90,
               target_width=120, name='PixelEBM', device=None, seed=None):
    # obs_spec is (observation_spec dict, action_spec) like get_cloning_network.py:99-102
    ob_spec, act_in_spec = obs_spec if isinstance(obs_spec, tuple) else (obs_spec, action_spec)
    rgb_shape = tuple(ob_spec['rgb'].shape)
    self.name = name
    self.state_spec = ()
    self._encoder = get_encoder_network(encoder_network, target_height, target_width,
                                        3 * rgb_shape[0])
    self._value = get_value_network(value_network)
    self.target_height, self.target_width = int(target_height), int(target_width)
    self.act_dim = int(act_in_spec.shape[-1])
    self.engine = eng.PixelEngine(rgb_shape[0], rgb_shape[1], rgb_shape[2], rgb_shape[3],
                                  self.target_height, self.target_width, self.act_dim,
                                  self._value.width, self._value.num_blocks, device)
    self.device = self.engine.device
    g = torch.Generator(device='cpu')
    if seed is not None:
      g.manual_seed(int(seed))
    flat = torch.zeros(self.engine.param_count, dtype=torch.float32)
    for nm, off, shape in self.param_layout():
      n = math.prod(shape)
      if nm.startswith('conv') and nm.endswith('_k'):      # Keras glorot_uniform, zero bias
        lim = math.sqrt(6.0 / (9 * shape[2] + 9 * shape[3]))
        flat[off:off + n] = (torch.rand(n, generator=g) * 2 - 1) * lim
      elif not nm.startswith('conv'):                      # 'normal' = N(0, 0.05)
        flat[off:off + n] = torch.randn(n, generator=g) * 0.05
    self.params = flat.to(self.device).contiguous()
    self.engine.bind_params(self.params)
    self.built = True
    self.losses = []

  def param_layout(self):
    a = self.engine.arch
    out, off = [], 0

    def add(nm, shape):
      nonlocal off
      out.append((nm, off, tuple(shape)))
      off += math.prod(shape)
    cin = a.channels * a.seq_len
    for i, co in enumerate(_CONV):
      add('conv%d_k' % (i + 1), (3, 3, cin, co))
      add('conv%d_b' % (i + 1), (co,))
      cin = co
    w = a.value_width
    add('v_dense0_W', (256 + a.act_dim, w)); add('v_dense0_b', (w,))
    for j in range(a.value_blocks):
      add('v_blk%d_d0_W' % j, (w, w // 4)); add('v_blk%d_d0_b' % j, (w // 4,))
      add('v_blk%d_d1_W' % j, (w // 4, w // 4)); add('v_blk%d_d1_b' % j, (w // 4,))
      add('v_blk%d_d2_W' % j, (w // 4, w)); add('v_blk%d_d2_b' % j, (w,))
    add('v_dense1_W', (w,)); add('v_dense1_b', (1,))
    return out

  @property
  def variables(self):
    return [self.params]

  trainable_variables = variables

  def create_variables(self, *unused_args, **unused_kwargs):
    return None

  def set_params(self, flat):
    self.params.copy_(flat.to(self.device, torch.float32).reshape(-1))

  def params_changed(self):
    return None         # the kernels read the flat vector directly

  def _images(self, obs):
    rgb = obs['rgb'] if isinstance(obs, dict) else obs
    return rgb.to(self.device).contiguous()

  def encode(self, obs, training=False):
    """Embeds images (pixel_ebm.py:79-96) -> [B, 256]."""
    del training
    return self.engine.encode(self._images(obs))

  def __call__(self, inputs, training=False, step_type=(), network_state=(),
               observation_encoding=None):
    del step_type
    obs, act = inputs
    act = act.to(self.device, torch.float32).contiguous()
    if observation_encoding is None:
      enc = self.encode(obs, training)
      n = act.shape[0] // enc.shape[0]                      # pixel_ebm.py:111-114
      return self.engine.energy(enc, act, n), network_state
    enc_t = observation_encoding.to(self.device, torch.float32).contiguous()
    if enc_t.shape[0] != act.shape[0]:
      raise ValueError('observation_encoding must be tiled to the actions')
    return self.engine.energy(enc_t, act, 1), network_state

  call = __call__
