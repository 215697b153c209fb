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


# ---- pushing_pixels/pixel_ebm_langevin.gin: dE/dact, Langevin and the gradient penalty on the
# late-fusion value head -----------------------------------------------------------------------
def _fro(got, want):
  got, want = got.detach().double().cpu(), want.detach().double().cpu()
  assert got.shape == want.shape, (got.shape, want.shape)
  return float((got - want).norm()) / (float(want.norm()) + 1e-30)


# Frobenius bounds: a ReLU mask that flips under operand rounding changes one row's gradient
# discontinuously, so dE/dact is bounded in aggregate rather than element by element
DACT_TOL = {'fp32': 1e-4, 'tf32': 2e-2}


def _scaled_params(spec, scale=4.0, seed=1):
  """N(0, 0.05) weights give |dE/dact| far below the penalty's margin; scaled weights make
  the penalty (and the clipped Langevin steps) active."""
  return opix.init_params(spec, seed=seed) * scale


def _engine_with(spec, flat, precision):
  e, _ = _engine(spec, precision=precision)
  dev = flat.cuda().contiguous()
  e.bind_params(dev)
  e._keep = dev
  return e


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('spec,b,n', [(SMALL, 3, 17), (WIDE, 2, 64)], ids=['small', 'wide'])
def test_value_energy_dact_matches_oracle_autograd(spec, b, n, precision):
  flat = _scaled_params(spec)
  e = _engine_with(spec, flat, precision)
  g = torch.Generator().manual_seed(6)
  enc = torch.rand(b, 256, generator=g)
  act = torch.rand(b * n, 2, generator=g) * 2 - 1
  a = act.double().clone().requires_grad_(True)
  want_e = opix.value(flat.double(), spec, torch.repeat_interleave(enc, n, 0).double(), a)
  (want_g,) = torch.autograd.grad(want_e.sum(), a)
  got_e, got_g = e.energy_dact(enc.cuda(), act.cuda(), n)
  assert e.kernel_status() == 0
  _close(got_e, want_e, ACT_TOL[precision], 'energy')
  assert float(want_g.abs().max()) > 1e-3
  drift = 0.0
  if precision == 'tf32':
    # Operand rounding flips the ReLU mask of units that sit within ~1e-3 of zero, and in the
    # 4x-narrower bottleneck layers one flip moves a row's dE/dact by 10-20 %.  The oracle's
    # tf32 emulation flips the same units: the kernels must reproduce it closely and stay
    # within 1.5x its distance from the exact gradient.
    from oracle import tf32_emul
    a2 = act.double().clone().requires_grad_(True)
    (em_g,) = torch.autograd.grad(
        tf32_emul.pixel_value(flat.double(), spec, enc.double(), a2).sum(), a2)
    drift = _fro(em_g, want_g)
    assert _fro(got_g, em_g) < 5e-3 + 0.25 * drift, (_fro(got_g, em_g), drift)
  assert _fro(got_g, want_g) < DACT_TOL[precision] + 1.5 * drift, (_fro(got_g, want_g), drift)
  # the same rows through the tiled-encoding form of mcmc.gradient_wrt_act (on the tf32
  # path the tiled encoding's projection is a tensor-core product of its own)
  got_e1, got_g1 = e.energy_dact(torch.repeat_interleave(enc, n, 0).cuda().contiguous(),
                                 act.cuda(), 1)
  assert _fro(got_g1, want_g) < DACT_TOL[precision] + 2.5 * drift, (_fro(got_g1, want_g), drift)
  _close(got_e1, want_e, ACT_TOL[precision], 'energy (tiled)')


@pytest.mark.parametrize('precision', PRECISIONS)
def test_pixel_langevin_chain_matches_oracle(precision):
  spec, b, n, k = SMALL, 2, 24, 6
  flat = _scaled_params(spec)
  e = _engine_with(spec, flat, precision)
  from ibc_b200 import engine as eng
  g = torch.Generator().manual_seed(7)
  enc = torch.rand(b, 256, generator=g)
  act0 = torch.rand(b * n, 2, generator=g) * 2.2 - 1.1
  normals = torch.randn(k, b * n, 2, generator=g)
  mn, mx = torch.full((2,), -1.1), torch.full((2,), 1.1)
  enc_t = torch.repeat_interleave(enc, n, 0).double()

  def energy_fn(_obs, a):
    return opix.value(flat.double(), spec, enc_t, a)

  want_a, want_chain = omcmc.langevin_actions_given_obs(
      energy_fn, None, act0.double(), mn.double(), mx.double(), normals.double(),
      num_iterations=k, sampler_stepsize_init=0.5, delta_action_clip=0.5, noise_scale=0.5,
      return_chain=True)
  cfg = eng.langevin_cfg(k, sampler_stepsize_init=0.5, delta_action_clip=0.5, noise_scale=0.5)
  actions = act0.cuda().clone()
  ca, ce, cg = e.langevin(enc.cuda(), actions, n, mn.cuda(), mx.cuda(), cfg,
                          normals.cuda().contiguous(), 0, return_chain=True)
  assert e.kernel_status() == 0
  # step 0 is evaluated at identical actions
  _close(ce[0], want_chain.energies[0], ACT_TOL[precision], 'chain energies[0]')
  assert _fro(cg[0], want_chain.grad_norms[0]) < DACT_TOL[precision]
  tol = {'fp32': 1e-4, 'tf32': 3e-2}[precision]
  assert _fro(actions, want_a) < tol
  assert _fro(ca, want_chain.actions) < tol
  assert _fro(ce, want_chain.energies) < tol
  assert torch.equal(ca[-1], actions)
  # device-generated noise: runs, stays in bounds, repeatable for a seed
  a1, a2 = act0.cuda().clone(), act0.cuda().clone()
  e.langevin(enc.cuda(), a1, n, mn.cuda(), mx.cuda(), cfg, None, 11)
  e.langevin(enc.cuda(), a2, n, mn.cuda(), mx.cuda(), cfg, None, 11)
  assert torch.equal(a1, a2) and bool((a1.abs() <= 1.1 + 1e-6).all())
  assert not torch.equal(a1, actions)


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('spec,b,n,norm', [(SMALL, 4, 8, 'inf'), (WIDE, 3, 8, 'inf'),
                                           (SMALL, 3, 5, '2'), (SMALL, 3, 5, '1')],
                         ids=['small-inf', 'wide-inf', 'small-l2', 'small-l1'])
def test_pixel_grad_penalty_matches_oracle_double_backward(spec, b, n, norm, precision):
  """ibc_pixel_train_grads with add_grad_penalty: the penalty term and its parameter
  gradients (GPU gradients with the penalty minus without) against the oracle's autograd
  double backward through PixelEBM.call (encode + tile + value head)."""
  from ibc_b200 import engine as eng
  flat = _scaled_params(spec)
  e = _engine_with(spec, flat, precision)
  img = _images(spec, b, seed=8)
  g = torch.Generator().manual_seed(9)
  act = torch.rand(b * (n + 1), 2, generator=g) * 2 - 1
  gb = 2 * b
  theta = flat.double().clone().requires_grad_(True)
  imgs = img.double() / 255.0

  def energy_fn(_obs, a):
    return opix.energy(theta, spec, imgs, a)

  # margin at the median norm: about half of the rows are penalised
  a0 = act.double().clone().requires_grad_(True)
  (g0,) = torch.autograd.grad(energy_fn(None, a0).sum(), a0)
  margin = float(omcmc.compute_grad_norm(norm, g0).median())
  assert margin > 1e-3
  want_gl = olosses.grad_penalty(energy_fn, norm, b, None, act.double(), grad_margin=margin)
  assert float(want_gl.detach().max()) > 0.0
  (want,) = torch.autograd.grad(want_gl.sum() / gb, theta)
  cfg_gp = eng.loss_cfg(1.0, True, norm, margin)
  cfg_no = eng.loss_cfg(1.0, False)
  per0, _, grads0 = e.train_grads(img.cuda(), act.cuda(), n, cfg_no, gb)
  grads0 = grads0.clone()
  per1, gl, _, grads1 = e.train_grads(img.cuda(), act.cuda(), n, cfg_gp, gb, want_grad_loss=True)
  assert e.kernel_status() == 0
  tol = {'fp32': 2e-3, 'tf32': 3e-2}[precision]
  assert _fro(gl, want_gl) < tol
  assert _fro(per1 - per0, want_gl) < tol + 1e-3
  got = (grads1 - grads0).double().cpu()
  wn = float(want.norm())
  for name, off, shape in opix.param_layout(spec):
    cnt = math.prod(shape)
    w, gt = want[off:off + cnt], got[off:off + cnt]
    if float(w.norm()) < 1e-9 * wn:
      # encoder, biases, encoding projection: dE/dact sees them through ReLU masks only.
      # (grads1 - grads0 still carries the fp32 atomics' ordering noise of the InfoNCE part)
      assert float(gt.norm()) < 1e-3 * wn + 1e-4 * float(grads0[off:off + cnt].norm()), name
      continue
    rel = float((gt - w).norm()) / (float(w.norm()) + 1e-3 * wn)
    assert rel < tol, (name, rel)


def test_late_fusion_agent_with_langevin_negatives_and_grad_penalty():
  """pushing_pixels/pixel_ebm_langevin.gin surface: Langevin counter-examples on the value
  head, gradient penalty, chain metrics; Langevin (+ again) policy inference."""
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import specs
  from ibc_b200.ibc.agents import ibc_agent, ibc_policy
  from ibc_b200.ibc.train import get_agent
  from ibc_b200.networks.pixel_ebm import PixelEBM
  gin.clear_config()
  gin.parse_config('DenseResnetValue.width = 64\nDenseResnetValue.num_blocks = 1\n'
                   'langevin_actions_given_obs.late_fusion = True\n'
                   'langevin_actions_given_obs.num_iterations = 20')
  try:
    obs_spec = {'rgb': specs.TensorSpec((2, 40, 56, 3), dtype=torch.uint8)}
    act_spec = specs.BoundedTensorSpec((2,), minimum=-1.0, maximum=1.0)
    samp = specs.BoundedTensorSpec((2,), minimum=-1.1, maximum=1.1)
    net = PixelEBM((obs_spec, act_spec), specs.TensorSpec((1,)), 'ConvMaxpoolEncoder',
                   'DenseResnetValue', target_height=32, target_width=48, seed=0)
    agent = ibc_agent.ImplicitBCAgent(obs_spec, act_spec, samp, net, get_agent.Adam(1e-3),
                                      num_counter_examples=8, fraction_langevin_samples=1.0,
                                      late_fusion=True, add_grad_penalty=True, compute_mse=True,
                                      run_full_chain_under_gradient=True)
    g = torch.Generator().manual_seed(0)
    level = torch.randint(0, 200, (8, 1, 1, 1, 3), generator=g)
    img = (level + torch.randint(0, 56, (8, 2, 40, 56, 3), generator=g)).to(torch.uint8)
    act = (level[:, 0, 0, 0, :2].float() / 200.0) * 2 - 1
    torch.manual_seed(0)
    losses = [float(agent.train((({'rgb': img}, act), ())).loss) for _ in range(60)]
    assert all(math.isfinite(x) for x in losses)
    assert min(losses[-10:]) < losses[0] - 0.05, (losses[0], losses[-10:])
    d = agent.last_losses_dict
    for key in ('grad_loss', 'overall_energies_avg', 'final_grad_norms_avg',
                'mse_counter_examples'):
      assert key in d and math.isfinite(float(d[key])), key
    ev = agent.get_eval_loss((({'rgb': img}, act), ()))
    assert math.isfinite(float(ev['ebm_total_loss']))
    # mcmc.gradient_wrt_act on the network: tiled encoding and self-encoding forms agree
    from ibc_b200.ibc.agents import mcmc
    acts = torch.rand(8 * 3, 2).cuda() * 2 - 1
    enc = net.encode({'rgb': img.cuda()})
    g1, e1 = mcmc.gradient_wrt_act(net, {'rgb': img.cuda()}, acts, False, (), (), False,
                                   specs.tile_batch(enc, 3))
    g2, e2 = mcmc.gradient_wrt_act(net, {'rgb': img.cuda()}, acts, False, (), (), False)
    assert torch.allclose(e1, e2, atol=2e-3) and torch.allclose(g1, g2, atol=2e-3)
    pol = ibc_policy.GreedyPolicy(ibc_policy.IbcPolicy(
        obs_spec, act_spec, samp, net, num_action_samples=64, use_dfo=False, use_langevin=True,
        optimize_again=True, late_fusion=True))
    a = pol.action(specs.restart({'rgb': img[:1]})).action
    assert tuple(a.shape) == (1, 2) and bool((a.abs() <= 1.0).all())
  finally:
    gin.clear_config()


@pytest.mark.parametrize('precision', PRECISIONS)
def test_train_begin_finish_equals_train_grads(precision):
  """ibc_pixel_train_begin / _finish with sampler calls in between == ibc_pixel_train_grads."""
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  spec, b, n = SMALL, 4, 8
  flat = _scaled_params(spec)
  e = _engine_with(spec, flat, precision)
  img = _images(spec, b, seed=12).cuda()
  g = torch.Generator().manual_seed(13)
  act = (torch.rand(b * (n + 1), 2, generator=g) * 2 - 1).cuda()
  cfg = eng.loss_cfg(1.0, True, 'inf', 0.05)
  per0, gl0, en0, gr0 = e.train_grads(img, act, n, cfg, b, want_energies=True, want_grad_loss=True)
  gr0 = gr0.clone()
  enc_ref = e.encode(img)
  enc = e.train_begin(img, n)
  assert torch.equal(enc, enc_ref)
  # sampler traffic on the handle between the two halves
  mn, mx = torch.full((2,), -1.1).cuda(), torch.full((2,), 1.1).cuda()
  neg = (torch.rand(b * n, 2, generator=g) * 2 - 1).cuda()
  e.langevin(enc, neg, n, mn, mx, eng.langevin_cfg(5), None, 3)
  e.dfo(enc, neg.clone(), n, mn, mx, seed=4)
  e.energy_dact(enc, neg, n)
  per1, gl1, en1, gr1 = e.train_finish(act, cfg, b, want_energies=True)
  assert e.kernel_status() == 0
  assert torch.equal(en0, en1) and torch.equal(per0, per1) and torch.equal(gl0, gl1)
  assert _fro(gr1, gr0) < 1e-5        # split-K atomics order only
  # a discarded step fails loudly
  e.train_begin(img, n)
  e.encode(img)
  with pytest.raises(ValueError):      # IBC_ERR_INVALID from the library
    e.train_finish(act, cfg, b)
  with pytest.raises(_lib.IbcError):   # the wrapper's own guard
    e.train_finish(act, cfg, b)
