"""GPU parity tests of the late-fusion PixelEBM against the oracle
(oracle/pixel_ebm.py) on identical inputs, in both precisions of the handle.
IBC_PRECISION_FP32 (fp32 FMA kernels vs the fp64 oracle): |d| <= 1e-4 * (|x| + rms(x))
for activations / energies, 2e-3 relative Frobenius error per parameter tensor for
gradients (long fp32 reductions over B*H*W rows).  IBC_PRECISION_TF32 (tcgen05
kind::tf32 operands rounded to nearest, fp32 accumulation): 3e-3 of the
same form for activations / energies; gradients are bounded against the oracle's tf32
emulation (see test_train_grads_match_oracle_autograd)."""
import math

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import mcmc as omcmc
from oracle import pixel_ebm as opix

pytestmark = pytest.mark.gpu

SMALL = opix.PixelSpec(seq_len=2, height=40, width=56, channels=3, target_height=32,
                       target_width=48, act_dim=2, value_width=64, value_blocks=2)


def _close(got, want, tol, name=''):
  got, want = got.detach().double().cpu(), want.detach().double().cpu()
  assert got.shape == want.shape, (name, got.shape, want.shape)
  rms = float(want.pow(2).mean().sqrt())
  worst = float(((got - want).abs() / (tol * (want.abs() + rms) + 1e-30)).max())
  assert worst <= 1.0, '%s: err/bound %.3g (rms %.3g)' % (name, worst, rms)


# value head 128 wide: the width from which the fused action-projection kernels of the tf32
# path (act_rows_kernel / act_rows_bwd_kernel, 128-column blocks) take over
WIDE = opix.PixelSpec(seq_len=2, height=40, width=56, channels=3, target_height=32,
                      target_width=48, act_dim=2, value_width=128, value_blocks=1)

PRECISIONS = ['fp32', 'tf32']
ACT_TOL = {'fp32': 1e-4, 'tf32': 3e-3}
GRAD_TOL = {'fp32': 2e-3, 'tf32': 5e-3}


def _engine(spec, seed=1, precision='fp32'):
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  e = eng.PixelEngine(spec.seq_len, spec.height, spec.width, spec.channels, spec.target_height,
                      spec.target_width, spec.act_dim, spec.value_width, spec.value_blocks)
  e.set_precision(_lib.PRECISION_TF32 if precision == 'tf32' else _lib.PRECISION_FP32)
  flat = opix.init_params(spec, seed=seed)
  assert e.param_count == opix.param_count(spec)
  dev = flat.cuda().contiguous()
  e.bind_params(dev)
  e._keep = dev
  return e, flat


def _images(spec, b, seed=0, u8=True):
  g = torch.Generator().manual_seed(seed)
  img = torch.randint(0, 256, (b, spec.seq_len, spec.height, spec.width, spec.channels),
                      generator=g, dtype=torch.uint8)
  return img if u8 else img.float() / 255.0


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('u8', [True, False])
def test_encode_matches_oracle(u8, precision):
  e, flat = _engine(SMALL, precision=precision)
  img = _images(SMALL, 3, u8=u8)
  got = e.encode(img.cuda())
  assert e.kernel_status() == 0
  want = opix.encode(flat.double(), SMALL, img.double() / 255.0 if u8 else img.double())
  _close(got, want, ACT_TOL[precision], 'encode')


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('b,n', [(3, 17), (2, 257)])
def test_value_energy_matches_oracle(b, n, precision):
  e, flat = _engine(SMALL, precision=precision)
  g = torch.Generator().manual_seed(2)
  enc = torch.rand(b, 256, generator=g)
  act = torch.rand(b * n, 2, generator=g) * 2 - 1
  got = e.energy(enc.cuda(), act.cuda(), n)
  want = opix.value(flat.double(), SMALL, torch.repeat_interleave(enc, n, 0).double(), act.double())
  _close(got, want, ACT_TOL[precision], 'value energy')


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('spec,b,n', [(SMALL, 4, 7), (SMALL, 6, 40), (WIDE, 5, 33)],
                         ids=['small-4-7', 'small-6-40', 'wide-5-33'])
def test_train_grads_match_oracle_autograd(spec, b, n, precision):
  from ibc_b200 import engine as eng
  SMALL = spec          # noqa: N806  (the body below is written against one spec)
  e, flat = _engine(SMALL, precision=precision)
  img = _images(SMALL, b, seed=3)
  g = torch.Generator().manual_seed(4)
  act = torch.rand(b * (n + 1), 2, generator=g) * 2 - 1
  theta = flat.double().clone().requires_grad_(True)
  pred = opix.energy(theta, SMALL, img.double() / 255.0, act.double()).reshape(b, n + 1)
  per = olosses.info_nce(pred, b, n, 1.0)
  gb = 2 * b
  (want,) = torch.autograd.grad(per.sum() / gb, theta, retain_graph=True)
  dpred_l1 = float(torch.autograd.grad(per.sum() / gb, pred)[0].abs().sum())
  per_g, en_g, grads = e.train_grads(img.cuda(), act.cuda(), n, eng.loss_cfg(1.0, False), gb,
                                     want_energies=True)
  assert e.kernel_status() == 0
  _close(en_g, pred.reshape(-1), ACT_TOL[precision], 'energies')
  _close(per_g, per, ACT_TOL[precision], 'per-example loss')
  want_em = None
  if precision == 'tf32':
    # The encoding gradient is a cancelling sum over an observation's N + 1 rows, so on
    # small batches tf32 operand rounding alone moves the conv gradients by 1-8 %.  The
    # oracle's tf32 emulation (operands of every product rounded, fp64 accumulation) says
    # how far a correct kind::tf32 implementation drifts; the kernels must stay within
    # 1e-2 of it and within 1.5x its distance from the exact gradient.
    from oracle import tf32_emul
    theta2 = flat.double().clone().requires_grad_(True)
    pred2 = tf32_emul.pixel_energy(theta2, SMALL, img.double() / 255.0,
                                   act.double()).reshape(b, n + 1)
    (want_em,) = torch.autograd.grad(olosses.info_nce(pred2, b, n, 1.0).sum() / gb, theta2)
  grms = float(want.pow(2).mean().sqrt())
  for name, off, shape in opix.param_layout(SMALL):
    cnt = math.prod(shape)
    w = want[off:off + cnt]
    got = grads[off:off + cnt].double().cpu()
    denom = float(w.norm()) + 1e-2 * grms * math.sqrt(cnt)
    if name == 'v_dense1_b':
      denom = dpred_l1     # sum of softmax gradients cancels to ~0: compare to the terms
    rel = float((got - w).norm()) / denom
    if want_em is None:
      assert rel < GRAD_TOL[precision], (name, rel)
    else:
      em = want_em[off:off + cnt]
      rel_em = float((em - w).norm()) / denom
      assert rel < GRAD_TOL[precision] + 1.5 * rel_em, (name, rel, rel_em)
      assert float((got - em).norm()) / denom < 1e-2 + rel_em, (name, float((got - em).norm()) / denom)


def test_pixel_dfo_pipeline_bit_exact_given_gpu_energies():
  e, _ = _engine(SMALL)
  b, n, iters = 2, 256, 3
  img = _images(SMALL, b, seed=5)
  enc = e.encode(img.cuda())
  rs = np.random.RandomState(1)
  uniforms = rs.rand(iters, b, n)
  normals = torch.from_numpy(rs.randn(iters - 1, b * n, 2).astype(np.float32))
  act0 = torch.from_numpy(rs.rand(b * n, 2).astype(np.float32)) * 2 - 1
  mn, mx = torch.full((2,), -1.1), torch.full((2,), 1.1)

  def gpu_energy(_obs, actions):
    return e.energy(enc, actions.float().cuda().contiguous(), n).cpu()

  want_probs, want_actions = omcmc.iterative_dfo(gpu_energy, b, None, act0, n, mn, mx, uniforms,
                                                 normals)
  actions = act0.cuda().clone()
  probs = e.dfo(enc, actions, n, mn.cuda(), mx.cuda(), 1.0, iters, 0.33,
                torch.from_numpy(uniforms).cuda(), normals.cuda())
  assert torch.equal(actions.cpu(), want_actions)
  _close(probs, want_probs, 1e-5, 'dfo probs')


def test_network_agent_and_policy_surface():
  """PixelEBM / late-fusion ImplicitBCAgent / IbcPolicy with the reference's
  call protocol: a few training steps reduce the loss; the DFO policy returns a
  bounded action."""
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import specs
  from ibc_b200.ibc.agents import ibc_agent, ibc_policy
  from ibc_b200.ibc.train import get_agent
  from ibc_b200.networks.pixel_ebm import PixelEBM
  gin.clear_config()
  gin.parse_config('DenseResnetValue.width = 64\nDenseResnetValue.num_blocks = 1')
  try:
    obs_spec = {'rgb': specs.TensorSpec((2, 40, 56, 3), dtype=torch.uint8)}
    act_spec = specs.BoundedTensorSpec((2,), minimum=-1.0, maximum=1.0)
    samp = specs.BoundedTensorSpec((2,), minimum=-1.1, maximum=1.1)
    net = PixelEBM((obs_spec, act_spec), specs.TensorSpec((1,)), 'ConvMaxpoolEncoder',
                   'DenseResnetValue', target_height=32, target_width=48, seed=0)
    with pytest.raises(ValueError):
      PixelEBM((obs_spec, act_spec), specs.TensorSpec((1,)), 'Nope', 'DenseResnetValue')
    opt = get_agent.Adam(1e-3)
    agent = ibc_agent.ImplicitBCAgent(obs_spec, act_spec, samp, net, opt, num_counter_examples=16,
                                      fraction_langevin_samples=0.0, late_fusion=True,
                                      add_grad_penalty=False, compute_mse=True)
    g = torch.Generator().manual_seed(0)
    # the action is readable from the mean brightness of two colour channels
    level = torch.randint(0, 200, (8, 1, 1, 1, 3), generator=g)
    img = (level + torch.randint(0, 56, (8, 2, 40, 56, 3), generator=g)).to(torch.uint8)
    act = (level[:, 0, 0, 0, :2].float() / 200.0) * 2 - 1
    torch.manual_seed(0)                      # counter-example draws are seeded from torch's RNG
    losses = [float(agent.train((({'rgb': img}, act), ())).loss) for _ in range(80)]
    assert min(losses[-10:]) < losses[0] - 0.1, (losses[0], losses[-10:])
    # network call protocol
    enc = net.encode({'rgb': img.cuda()}, training=False)
    assert tuple(enc.shape) == (8, 256)
    acts = torch.rand(8 * 5, 2).cuda()
    e1, _ = net(({'rgb': img.cuda()}, acts), training=False)
    e2, _ = net(({'rgb': img.cuda()}, acts), training=False,
                observation_encoding=specs.tile_batch(enc, 5))
    # tf32 default: the tiled encoding's projection is a 40-row block on the tensor-core
    # kernel, the un-tiled one an 8-row block on the fp32 kernel
    assert torch.allclose(e1, e2, atol=2e-3)
    pol = ibc_policy.GreedyPolicy(ibc_policy.IbcPolicy(
        obs_spec, act_spec, samp, net, num_action_samples=256, use_dfo=True, use_langevin=False,
        late_fusion=True))
    a = pol.action(specs.restart({'rgb': img[:1]})).action
    assert tuple(a.shape) == (1, 2) and bool((a.abs() <= 1.0).all())
  finally:
    gin.clear_config()


@pytest.mark.parametrize('precision', PRECISIONS)
def test_full_size_frames_run(precision):
  """pushing_pixels/pixel_ebm_best.gin shapes: 2 x 240x320x3 -> 180x240, Wv=1024, 1 block."""
  spec = opix.PixelSpec()
  e, flat = _engine(spec, precision=precision)
  img = _images(spec, 2, seed=9)
  enc = e.encode(img.cuda())
  assert e.kernel_status() == 0
  want = opix.encode(flat.double(), spec, img.double() / 255.0)
  _close(enc, want, 2 * ACT_TOL[precision], 'full-size encode')
