"""Pins the CPU oracle: the reference's behavioural tests
(ibc/agents/mcmc_test.py) re-run against it, the hand-derived known answers of
SURVEY.md Appendix B, and autograd cross-checks of the manual formulas."""
import math

import numpy as np
import pytest
import torch

from oracle import agent as oagent
from oracle import losses as olosses
from oracle import mcmc as omcmc
from oracle import mlp_ebm as omlp
from oracle import pixel_ebm as opix
from oracle import resample_c


def mock_energy(mean=(0.3, 0.4), scale=1e2):
  """mcmc_test.py:67-81."""
  mean_t = torch.tensor(mean, dtype=torch.float32)

  def fn(obs, actions):
    return -(torch.linalg.norm(actions - mean_t, dim=1) * scale) ** 2
  return fn, mean_t


# ------------------------------ resampling --------------------------------
def test_exp64_matches_libm_and_c():
  xs = -np.abs(np.random.RandomState(0).randn(2000)) * 30
  xs = np.concatenate([xs, [0.0, -1e-300, -707.9, -708.0, -709.0, -1000.0]])
  ours = omcmc.exp64(xs)
  ref = np.exp(xs)
  ok = xs >= -708.0
  rel = np.abs(ours[ok] - ref[ok]) / ref[ok]
  assert rel.max() < 4e-16
  assert np.all(ours[~ok] == 0.0)
  c = np.array([resample_c.exp64(x) for x in xs])
  assert np.array_equal(c, ours)      # bit-identical NumPy vs C


def test_bincount_shapes():
  # mcmc_test.py:30-43
  rs = np.random.RandomState(0)
  for b in [1, 2]:
    probs = rs.rand(b, 256)
    probs /= probs.sum(1, keepdims=True)
    u = rs.rand(b, 128)
    counts = omcmc.categorical_bincount(128, np.log(probs), 256, u)
    assert counts.shape == (b, 256)


def test_bincount_correct():
  # mcmc_test.py:45-57
  eps = 1e-6
  probs = np.array([[eps, eps, 1. - eps], [eps, 1. - eps, eps]])
  u = np.random.RandomState(1).rand(2, 5)
  counts = omcmc.categorical_bincount(5, np.log(probs), 3, u)
  assert counts[0].sum() == 5 and counts[1].sum() == 5
  assert np.argmax(counts[0]) == 2 and np.argmax(counts[1]) == 1


def test_resample_known_answer():
  # Appendix B.4
  logits = np.array([[0.0, math.log(3.0)]], dtype=np.float32)
  u = np.array([[0.10, 0.20, 0.30, 0.90]])
  draws = omcmc.multinomial_draws(logits, u)
  assert draws.tolist() == [[0, 0, 1, 1]]
  assert omcmc.categorical_bincount(4, logits, 2, u).tolist() == [[2, 2]]
  assert omcmc.resample_indices(logits, u).tolist() == [0, 0, 1, 1]
  # -inf logit: class skipped
  logits = np.array([[-np.inf, 0.0, 0.0]], dtype=np.float32)
  u = np.array([[0.0, 0.25, 0.75, 0.999]])
  assert omcmc.multinomial_draws(logits, u).tolist() == [[1, 1, 2, 2]]


def test_resample_numpy_equals_c():
  rs = np.random.RandomState(2)
  for b, n, scale in [(3, 257, 1.0), (2, 2048, 30.0), (1, 16384, 5.0)]:
    logits = (rs.randn(b, n) * scale).astype(np.float32)
    logits[0, 3] = -np.inf
    u = rs.rand(b, n)
    draws_np = omcmc.multinomial_draws(logits, u)
    counts_np = omcmc.categorical_bincount(n, logits, n, u)
    rows_np = omcmc.resample_indices(logits, u)
    draws_c, counts_c, rows_c = resample_c.resample(logits, u)
    assert np.array_equal(draws_np, draws_c)
    assert np.array_equal(counts_np, counts_c)
    assert np.array_equal(rows_np, rows_c)
    assert np.all(np.diff(rows_c.reshape(b, n), axis=1) >= 0)    # sortedness


# ------------------------------ DFO / Langevin ----------------------------
def test_dfo_correct_on_mock_energy():
  # mcmc_test.py:117-144: B=2, N=2048, 10 iterations, std 0.1 -> within 0.01
  fn, mean = mock_energy()
  rs = np.random.RandomState(3)
  b, n = 2, 2048
  init = torch.tensor(rs.rand(b * n, 2), dtype=torch.float32)
  uniforms = rs.rand(10, b, n)
  normals = torch.tensor(rs.randn(9, b * n, 2), dtype=torch.float32)
  probs, actions = omcmc.iterative_dfo(
      fn, b, (), init, n, torch.zeros(2), torch.ones(2), uniforms, normals,
      num_iterations=10, iteration_std=1e-1)
  assert probs.shape == (b * n,) and actions.shape == init.shape
  assert torch.linalg.norm(actions[int(torch.argmax(probs))] - mean) < 0.01


def test_dfo_probs_actions_misalignment():
  # Appendix B.5: probs are those of the PRE-resample rows.
  x = torch.tensor([[0.1, 0.1], [0.7, 0.7]])

  def fn(obs, a):
    return torch.where(a[:, 0] > 0.5, torch.tensor(20.0), torch.tensor(0.0))
  u = np.array([[[0.3, 0.6]]])
  probs, actions = omcmc.iterative_dfo(fn, 1, (), x, 2, torch.zeros(2),
                                       torch.ones(2), u, None,
                                       num_iterations=1)
  assert torch.allclose(actions, x[[1, 1]])
  assert probs[1] > 0.999 and probs[0] < 1e-8


def test_langevin_correct_on_mock_energy():
  # mcmc_test.py:146-170: N=128, 25 iterations -> actions[0] within 0.1
  fn, mean = mock_energy()
  rs = np.random.RandomState(4)
  b, n = 2, 128
  init = torch.tensor(rs.rand(b * n, 2), dtype=torch.float32)
  normals = torch.tensor(rs.randn(25, b * n, 2), dtype=torch.float32)
  actions = omcmc.langevin_actions_given_obs(
      fn, (), init, torch.zeros(2), torch.ones(2), normals, num_iterations=25)
  assert torch.linalg.norm(actions[0] - mean) < 0.1


def test_langevin_single_step_known_answer():
  # Appendix B.3
  fn, _ = mock_energy()
  a = torch.tensor([[0.5, 0.5]])
  new_a, e, gn = omcmc.langevin_step(
      fn, (), a, torch.zeros(1, 2), 1.0, None, 0.1, 0.1, torch.zeros(2),
      torch.ones(2), 'inf')
  assert torch.allclose(new_a, torch.tensor([[0.45, 0.45]]))
  assert abs(float(gn[0]) - 4000.0) < 1e-2
  assert abs(float(e[0]) + 1e4 * (0.04 + 0.01)) < 1e-1


def test_polynomial_schedule_known_answers():
  # Appendix B.2
  t = omcmc.stepsize_table(100)
  assert t[0] == 0.1
  assert abs(t[1] - 0.0979902020) < 1e-9
  assert abs(t[50] - 0.0245050505) < 1e-9
  assert abs(t[99] - 1e-5) < 1e-15
  t2 = omcmc.stepsize_table(100, init=1e-5, final=1e-5)
  assert all(abs(x - 1e-5) < 1e-20 for x in t2)


def test_chain_data_layout():
  fn, _ = mock_energy()
  rs = np.random.RandomState(5)
  init = torch.tensor(rs.rand(8, 2), dtype=torch.float32)
  normals = torch.tensor(rs.randn(4, 8, 2), dtype=torch.float32)
  final, chain = omcmc.langevin_actions_given_obs(
      fn, (), init, torch.zeros(2), torch.ones(2), normals, num_iterations=4,
      return_chain=True)
  assert chain.actions.shape == (4, 8, 2)
  assert torch.equal(chain.actions[-1], final)           # post-step actions
  assert torch.allclose(chain.energies[0], fn((), init))  # pre-step energies


# ------------------------------ losses -------------------------------------
def test_info_nce_known_answer():
  # Appendix B.1: equal predictions, N=8 -> 2.19721344 (Keras epsilon form)
  pred = torch.zeros(3, 9, dtype=torch.float64)
  loss = olosses.info_nce(pred, 3, 8)
  expect = math.log(9) + 8 * 1e-7 * (math.log(1e-7) - math.log(1 / 9))
  assert torch.allclose(loss, torch.full((3,), expect, dtype=torch.float64),
                        atol=1e-12)
  assert abs(expect - 2.19721344) < 1e-8


def test_grad_penalty_linear_net():
  # Appendix B.8: E = a.w -> penalty max(0, |w|_inf - 1)^2, grad only at argmax
  w = torch.tensor([0.5, -3.0, 2.0], requires_grad=True)

  def fn(obs, a):
    return a @ w
  acts = torch.randn(6, 3)
  gl = olosses.grad_penalty(fn, 'inf', 2, None, acts)
  assert torch.allclose(gl, torch.full((2,), 4.0))
  (gw,) = torch.autograd.grad(gl.sum(), w)
  # d/dw of sum over 2 examples of mean over 3 rows of (|w|_inf-1)^2
  assert torch.allclose(gw, torch.tensor([0.0, -2 * 2.0 * 2, 0.0]))


def test_tile_batch_order():
  # Appendix B.6
  o = torch.tensor([[0.0], [1.0]])
  assert oagent.tile_batch(o, 3).reshape(-1).tolist() == [0, 0, 0, 1, 1, 1]


# ------------------------------ MLP EBM ------------------------------------
@pytest.mark.parametrize('spec', [
    omlp.MLPSpec(16, 2, 256, 2), omlp.MLPSpec(20, 2, 128, 16),
    omlp.MLPSpec(39, 28, 64, 4)])
def test_manual_dact_matches_autograd(spec):
  flat = omlp.init_params(spec, dtype=torch.float64) * 3
  g = torch.Generator().manual_seed(0)
  obs = torch.randn(37, spec.obs_dim, generator=g, dtype=torch.float64)
  act = torch.rand(37, spec.act_dim, generator=g, dtype=torch.float64) * 2 - 1
  e, d = omlp.energy_and_dact_manual(flat, spec, obs, act)
  a = act.clone().requires_grad_(True)
  e2 = omlp.energy(flat, spec, obs, a)
  (d2,) = torch.autograd.grad(e2.sum(), a)
  assert torch.allclose(e, e2, atol=1e-12)
  assert torch.allclose(d, d2, atol=1e-10)


def test_param_count_matches_survey():
  assert omlp.param_count(omlp.MLPSpec(20, 2, 128, 16)) == 267265
  assert omlp.param_count(omlp.MLPSpec(39, 28, 512, 8)) == 2136577
  assert omlp.param_count(omlp.MLPSpec(256, 32, 2048, 2)) == 8986625
  assert opix.param_count(opix.PixelSpec()) == 1246881


def test_spectral_norm():
  w = torch.randn(7, 5, dtype=torch.float64)
  u = torch.randn(1, 5, dtype=torch.float64)
  for _ in range(200):
    wbar, u = omlp.spectral_normalize(w, u)
  assert abs(float(torch.linalg.matrix_norm(wbar, ord=2)) - 1.0) < 1e-6


# ------------------------------ Adam ---------------------------------------
def test_adam_matches_torch_adam_form():
  # Keras form: theta -= lr_t * m / (sqrt(v) + eps) with eps OUTSIDE the
  # bias correction ("epsilon hat"); check the first step analytically.
  flat = torch.tensor([1.0, -2.0])
  grad = torch.tensor([0.5, -0.25])
  st = oagent.adam_init(flat)
  new = oagent.adam_update(flat, grad, st, base_lr=1e-3)
  lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
  m = 0.1 * grad
  v = 0.001 * grad * grad
  assert torch.allclose(new, flat - lr_t * m / (torch.sqrt(v) + 1e-7))
  assert st.iterations == 1
  assert abs(oagent.exponential_decay_lr(1e-3, 100) - 0.99e-3) < 1e-12


# ------------------------------ pixels -------------------------------------
def test_image_stacking_quirk():
  # Appendix B.7: raw reshape keeps memory order.
  spec = opix.PixelSpec(seq_len=2, height=2, width=2, channels=1,
                        target_height=2, target_width=2)
  img = torch.arange(8, dtype=torch.float32).reshape(1, 2, 2, 2, 1)
  out = opix.preprocess(img, spec)
  assert out.reshape(-1).tolist() == list(range(8))


def test_pixel_shapes():
  spec = opix.PixelSpec(height=48, width=64, target_height=32, target_width=48,
                        value_width=64)
  flat = opix.init_params(spec)
  img = torch.randint(0, 256, (2, 2, 48, 64, 3), dtype=torch.uint8)
  enc = opix.encode(flat, spec, img)
  assert enc.shape == (2, 256)
  e = opix.energy(flat, spec, img, torch.rand(6, 2))
  assert e.shape == (6,)


# ------------------------------ agent --------------------------------------
def test_train_step_decreases_loss_c1_small():
  spec = omlp.MLPSpec(16, 2, 32, 2)
  cfg = oagent.AgentConfig(num_counter_examples=16, fraction_langevin_samples=0.,
                           add_grad_penalty=False)
  g = torch.Generator().manual_seed(0)
  obs = torch.randn(32, 16, generator=g)
  act = torch.rand(32, 2, generator=g) * 2 - 1
  lo, hi = -1.1 * torch.ones(2), 1.1 * torch.ones(2)
  flat = omlp.init_params(spec)
  st = oagent.adam_init(flat)
  losses = []
  for it in range(30):
    u = torch.rand(32, 16, 2, generator=g)
    flat, total, grad, metrics = oagent.train_step(
        lambda th: (lambda o, a: omlp.energy(th, spec, o, a)), flat, st, cfg,
        obs, act, lo, hi, u, base_lr=1e-2)
    losses.append(float(total))
  assert losses[-1] < losses[0]


def test_langevin_train_loss_runs_and_grad_penalty_second_order():
  spec = omlp.MLPSpec(8, 3, 32, 2)
  cfg = oagent.AgentConfig(num_counter_examples=4, langevin_num_iterations=5)
  g = torch.Generator().manual_seed(1)
  obs = torch.randn(6, 8, generator=g)
  act = torch.rand(6, 3, generator=g) * 2 - 1
  lo, hi = -1.1 * torch.ones(3), 1.1 * torch.ones(3)
  flat = (omlp.init_params(spec) * 8).requires_grad_(True)
  u = torch.rand(6, 4, 3, generator=g)
  normals = torch.randn(5, 24, 3, generator=g)
  total, per, metrics = oagent.loss(
      lambda o, a: omlp.energy(flat, spec, o, a), cfg, obs, act, lo, hi, u,
      langevin_normals=normals)
  (grad,) = torch.autograd.grad(total, flat)
  assert torch.isfinite(grad).all() and 'grad_loss' in metrics
  assert 'final_grad_norms_avg' in metrics


# ---------------- independent restatements of the third-party semantics ----------------
# (the multinomial rule and the Keras KLD clipping were each pinned by one hand
# case.)  None of these share code with oracle/: plain Python loops, torch.nn modules, scipy.
def test_multinomial_rule_against_scalar_loop_and_frequencies():
  """tf.random.categorical on CPU (MultinomialFunctor): weights exp(logit - max) in double,
  running sum, first class whose running sum EXCEEDS u * total (std::upper_bound).  Restated
  here as a scalar Python loop with math.exp and checked draw by draw; then the empirical class
  frequencies of 200 000 draws against softmax(logits) (5 sigma per class)."""
  rs = np.random.RandomState(11)
  for n, scale in ((7, 1.0), (33, 4.0), (257, 12.0)):
    logits = (rs.randn(2, n) * scale).astype(np.float32)
    logits[1, n // 2] = -np.inf
    u = rs.rand(2, 64)
    got = omcmc.multinomial_draws(logits, u)
    for b in range(2):
      mx = max(float(x) for x in logits[b] if math.isfinite(x))
      run, cdf = 0.0, []
      for x in logits[b]:
        run += math.exp(float(x) - mx) if math.isfinite(x) else 0.0
        cdf.append(run)
      for k in range(u.shape[1]):
        target = u[b, k] * cdf[-1]
        idx = next((i for i, c in enumerate(cdf) if c > target), n - 1)
        assert got[b, k] == idx, (n, b, k)
  logits = np.array([[0.3, -1.2, 2.0, 0.0, -0.4]], dtype=np.float32)
  draws = 200_000
  counts = omcmc.categorical_bincount(draws, logits, 5, rs.rand(1, draws))[0]
  p = np.exp(logits[0].astype(np.float64))
  p /= p.sum()
  sigma = np.sqrt(draws * p * (1 - p))
  assert np.all(np.abs(counts - draws * p) < 5 * sigma), (counts, draws * p)


def test_info_nce_against_cross_entropy_and_scipy():
  """KLD(one_hot || softmax) is the cross entropy up to Keras' clipping of both arguments to
  [1e-7, 1]: against torch.nn.functional.cross_entropy where no probability is below 1e-7 (the
  clipped zeros of the label add N * 1e-7 * log(1e-7 / p_j) exactly), and term by term against
  scipy.special.rel_entr on the clipped arrays when predictions DO underflow the clip."""
  import scipy.special
  g = torch.Generator().manual_seed(4)
  for b, n, scale in ((5, 8, 1.0), (4, 256, 1.5), (3, 16, 40.0)):
    pred = torch.randn(b, n + 1, generator=g, dtype=torch.float64) * scale
    got = olosses.info_nce(pred, b, n)
    p = torch.softmax(pred, dim=-1).numpy()
    y = np.zeros_like(p)
    y[:, n] = 1.0
    want = scipy.special.rel_entr(np.clip(y, 1e-7, 1.0), np.clip(p, 1e-7, 1.0)).sum(axis=1)
    assert np.allclose(got.numpy(), want, rtol=1e-12, atol=1e-12)
    if scale <= 1.5:
      assert p.min() > 1e-7
      ce = torch.nn.functional.cross_entropy(pred, torch.full((b,), n), reduction='none').numpy()
      extra = (1e-7 * (math.log(1e-7) - np.log(p[:, :n]))).sum(axis=1)
      assert np.allclose(got.numpy(), ce + extra, rtol=1e-10, atol=1e-12)


def test_energy_against_independent_keras_style_modules():
  """MLPEBM with ResNetPreActivationLayer (networks/mlp_ebm.py:72-96, layers/resnet.py:131-153)
  rebuilt from torch.nn.Linear modules in Keras' own order -- Dense projection, then per block
  relu -> Dense -> relu -> Dense -> add -- with the oracle's flat parameters loaded as Keras
  kernels ([in, out] = Linear.weight.T); energies and autograd dE/dact must agree."""
  spec = omlp.MLPSpec(7, 3, 16, 4)
  flat = omlp.init_params(spec, seed=6, dtype=torch.float64)
  p = omlp.unpack(flat, spec)

  def dense(kernel, bias):
    lin = torch.nn.Linear(kernel.shape[0], kernel.shape[1] if kernel.dim() == 2 else 1).double()
    with torch.no_grad():
      lin.weight.copy_(kernel.reshape(kernel.shape[0], -1).T)
      lin.bias.copy_(bias.reshape(-1))
    return lin

  proj = dense(p['Wp'], p['bp'])
  blocks = [(dense(p['W1_%d' % j], p['b1_%d' % j]), dense(p['W2_%d' % j], p['b2_%d' % j]))
            for j in range(spec.num_blocks)]
  head = dense(p['we'], p['be'])
  g = torch.Generator().manual_seed(8)
  obs = torch.randn(9, 7, generator=g, dtype=torch.float64)
  act = torch.randn(9, 3, generator=g, dtype=torch.float64, requires_grad=True)
  x = proj(torch.cat([obs, act], dim=-1))
  for l1, l2 in blocks:
    y = l1(torch.relu(x))
    y = l2(torch.relu(y))
    x = x + y
  want = head(x).squeeze(-1)
  (dwant,) = torch.autograd.grad(want.sum(), act)
  got = omlp.energy(flat, spec, obs, act.detach())
  assert torch.allclose(got, want.detach(), rtol=1e-12, atol=1e-12)
  e2, dact = omlp.energy_and_dact_manual(flat, spec, obs, act.detach())
  assert torch.allclose(e2, want.detach(), rtol=1e-12, atol=1e-12)
  assert torch.allclose(dact, dwant, rtol=1e-10, atol=1e-12)


def test_spectral_norm_against_torch_power_iteration():
  """DenseSN (networks/layers/spectral_norm.py): one power iteration v = l2n(u W^T),
  u' = l2n(v W), sigma = v W u'^T, W / sigma.  After many iterations sigma must be the largest
  singular value (torch.linalg.svdvals) and the normalised kernel's spectral norm 1."""
  g = torch.Generator().manual_seed(2)
  w = torch.randn(24, 16, generator=g, dtype=torch.float64)
  u = torch.randn(1, 16, generator=g, dtype=torch.float64)
  for _ in range(300):
    w_bar, u = omlp.spectral_normalize(w, u)
  top = torch.linalg.svdvals(w)[0]
  assert torch.allclose(torch.linalg.svdvals(w_bar)[0], torch.ones((), dtype=torch.float64), atol=1e-9)
  assert torch.allclose(w / top, w_bar, atol=1e-9)


def test_cd_losses_known_answers():
  """ibc/losses/ebm_loss.py:51-66, 109-148 by hand: true action (0, 0); counter examples at squared
  distance 1 (outside the 0.2 threshold), 0.01 (inside), 0.25 (outside)."""
  pred = torch.tensor([[3.0, 5.0, 7.0, 1.0]], dtype=torch.float64)
  assert olosses.cd(pred).tolist() == [4.0]                          # mean(3, 5, 7) - 1
  counter = torch.tensor([[[1.0, 0.0], [0.1, 0.0], [0.5, 0.0]]], dtype=torch.float64)
  true = torch.zeros(1, 3, 2, dtype=torch.float64)
  loss, dbg = olosses.clipped_cd(pred, counter, true, soft=False)
  assert abs(float(loss) - 8.0 / 3.0) < 1e-12                        # ((3-1) + 0 + (7-1)) / 3
  assert abs(float(dbg['fraction_outside_threshold']) - 2.0 / 3.0) < 1e-12
  loss, dbg = olosses.clipped_cd(pred, counter, true, soft=True)
  soft = abs((1.0 - 5.0) - 0.1 ** 2) / 3.0                           # |(data - samp) - err| inside
  assert abs(float(loss) - (8.0 / 3.0 + soft)) < 1e-12
  assert abs(float(dbg['soft_loss']) - soft) < 1e-12 and abs(float(dbg['cd_loss']) - 8.0 / 3.0) < 1e-12
