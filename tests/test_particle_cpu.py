"""CPU tests of the offline-runnable task and the train_eval data path:
particle environment + scripted oracle + metrics (environments/particle/*.py), Chan
statistics and normaliser layers (ibc/train/stats.py, the cases of stats_test.py), the
sampling spec (get_sampling_spec.py) and the sequence windows (get_data.py)."""
import numpy as np
import pytest
import torch

from ibc_b200.environments.particle import HistoryWrapper, ParticleEnv, ParticleOracle
from ibc_b200.environments.particle.particle_metrics import ParticleMetrics
from ibc_b200.ibc import specs
from ibc_b200.ibc.data import in_memory
from ibc_b200.ibc.train import get_sampling_spec, stats


def test_particle_env_shapes_and_episode_length():
  env = ParticleEnv(n_dim=3, seed=1)
  obs = env.reset()
  assert list(obs) == ['pos_agent', 'vel_agent', 'pos_first_goal', 'pos_second_goal']
  assert all(v.shape == (3,) and v.dtype == np.float32 for v in obs.values())
  assert np.all(obs['vel_agent'] == 0)
  done, steps = False, 0
  while not done:
    obs, reward, done, _ = env.step(np.full(3, 0.5, np.float32))
    steps += 1
    assert reward in (0.0, 1.0)
  assert steps == env.n_steps == 50
  assert len(env.obs_log) == 1 + 50 * env.repeat_actions
  hidden = ParticleEnv(hide_velocity=True, seed=1)
  assert 'vel_agent' not in hidden.reset()


def test_particle_dynamics_known_answer():
  """One internal PD step (particle.py:175-191): x' = x + v dt, v' = v + (kp (a - x) - kv v) dt."""
  env = ParticleEnv(n_dim=2, seed=0, repeat_actions=1, dt=0.005)
  o0 = env.reset()
  a = np.array([1.0, 0.0], np.float32)
  o1, _, _, _ = env.step(a)
  np.testing.assert_allclose(o1['pos_agent'], o0['pos_agent'], rtol=0, atol=0)   # v0 = 0
  np.testing.assert_allclose(o1['vel_agent'], 10.0 * (a - o0['pos_agent']) * 0.005, rtol=1e-6)
  o2, _, _, _ = env.step(a)
  np.testing.assert_allclose(o2['pos_agent'], o1['pos_agent'] + o1['vel_agent'] * 0.005, rtol=1e-6)
  u = 10.0 * (a - o1['pos_agent']) - 5.0 * o1['vel_agent']
  np.testing.assert_allclose(o2['vel_agent'], o1['vel_agent'] + u * 0.005, rtol=1e-5)


def test_oracle_solves_the_task_and_metrics():
  env = ParticleEnv(seed=0)
  oracle = ParticleOracle(env)
  metrics = ParticleMetrics()
  for _ in range(10):
    obs = env.reset()
    oracle.reset()
    done = False
    while not done:
      obs, reward, done, _ = env.step(oracle.action(obs))
    assert reward == 1.0 and env.succeeded
    metrics.record_episode_end(env)
  res = metrics.result()
  assert res['AverageSuccessMetric'] == 1.0
  assert res['AverageFirstGoalDistance'] < 0.02 and res['AverageFinalSecondGoalDistance'] < 0.02
  # a policy that never moves fails
  obs = env.reset()
  done = False
  while not done:
    obs, reward, done, _ = env.step(obs['pos_agent'])
  assert reward == 0.0 and not env.succeeded


def test_history_wrapper_tiles_first_observation():
  env = HistoryWrapper(ParticleEnv(seed=3), history_length=2)
  obs = env.reset()
  assert obs['pos_agent'].shape == (2, 2)
  np.testing.assert_array_equal(obs['pos_agent'][0], obs['pos_agent'][1])
  prev = obs['pos_agent'][1].copy()
  obs, _, _, _ = env.step(np.array([0.9, 0.1], np.float32))
  np.testing.assert_array_equal(obs['pos_agent'][0], prev)
  assert not np.array_equal(obs['pos_agent'][0], obs['pos_agent'][1])


def test_windows_from_episodes():
  eps = in_memory.collect_particle_episodes(3, n_dim=2, seed=5)
  obs, act = in_memory.windows_from_episodes(eps, sequence_length=2)
  assert act.shape == (3 * 50, 2) and obs['pos_agent'].shape == (3 * 50, 2, 2)
  e_obs, e_act = eps[0]
  # first step: the first frame is replicated (dataset.py:122-125), action of step 0
  np.testing.assert_array_equal(obs['pos_agent'][0], e_obs['pos_agent'][[0, 0]])
  np.testing.assert_array_equal(act[0], e_act[0])
  np.testing.assert_array_equal(obs['pos_agent'][1], e_obs['pos_agent'][0:2])
  np.testing.assert_array_equal(act[1], e_act[1])            # action of the window's last step
  np.testing.assert_array_equal(obs['pos_agent'][49], e_obs['pos_agent'][48:50])
  # second episode starts with its own replicated first frame
  np.testing.assert_array_equal(obs['pos_agent'][50], eps[1][0]['pos_agent'][[0, 0]])
  obs3, _ = in_memory.windows_from_episodes(eps[:1], sequence_length=3)
  np.testing.assert_array_equal(obs3['pos_agent'][1], e_obs['pos_agent'][[0, 0, 1]])
  # the oracle's action is one of the two goals
  assert np.all(np.isclose(act[:, None, :], np.stack(
      [obs['pos_first_goal'][:, -1], obs['pos_second_goal'][:, -1]], 1)).all(-1).any(-1))


def test_chan_statistics_match_numpy_for_any_split():
  rng = np.random.RandomState(0)
  x = rng.randn(1000, 3, 4) * np.array([1., 5., 0.1, 2.]) + np.array([0., 3., -2., 10.])
  st = stats.ChanRunningStatistics(x[0, 0])
  for lo, hi in ((0, 1), (1, 7), (7, 500), (500, 1000)):
    st.update_running_statistics(x[lo:hi])
  flat = x.reshape(-1, 4)
  assert st.n == flat.shape[0]
  np.testing.assert_allclose(st.mean, flat.mean(0), rtol=1e-10)
  np.testing.assert_allclose(st.std, flat.std(0), rtol=1e-10)       # population std
  single = stats.ChanRunningStatistics(flat[0])
  for row in flat[:50]:
    single.update_running_statistics(row)
  np.testing.assert_allclose(single.variance, flat[:50].var(0), rtol=1e-10)


def test_dataset_statistics_and_layers():
  """stats_test.py: flat / nested statistics, min-max actions, norm -> denorm round trip."""
  rng = np.random.RandomState(1)
  obs = {'a': rng.rand(400, 2, 3).astype(np.float32) * 4 + 1,
         'b': rng.randn(400, 2, 5).astype(np.float32)}
  act = (rng.rand(400, 2).astype(np.float32) - 0.3) * 3
  on, an, adn, amin, amax = stats.compute_dataset_statistics(
      obs, act, num_samples=300, nested_obs=True, min_max_actions=True)
  np.testing.assert_allclose(amin, act[:300].min(0))
  np.testing.assert_allclose(amax, act[:300].max(0))
  normed = on({k: torch.as_tensor(v[:300]) for k, v in obs.items()})
  for k in obs:
    flat = normed[k].reshape(-1, normed[k].shape[-1]).numpy()
    np.testing.assert_allclose(flat.mean(0), 0, atol=1e-4)
    np.testing.assert_allclose(flat.std(0), 1, atol=1e-4)
  na = an(torch.as_tensor(act[:300]))
  assert float(na.min()) == pytest.approx(-1.0, abs=1e-6)
  assert float(na.max()) == pytest.approx(1.0, abs=1e-6)
  np.testing.assert_allclose(adn(na).numpy(), act[:300], atol=1e-6)
  # z-scored actions and the flat-observation error of stats_test.py:58-62
  on2, an2, adn2, _, _ = stats.compute_dataset_statistics(obs['a'], act, num_samples=400,
                                                          nested_obs=False)
  z = an2(torch.as_tensor(act))
  np.testing.assert_allclose(z.numpy().mean(0), 0, atol=1e-5)
  np.testing.assert_allclose(adn2(z).numpy(), act, atol=1e-5)
  with pytest.raises(ValueError, match='too many'):
    stats.compute_dataset_statistics(obs, act, num_samples=10, nested_obs=False)
  # constant feature: std clamps to eps instead of dividing by zero
  const = stats.StdNormalizationLayer(mean=np.zeros(2), std=np.zeros(2))
  assert torch.isfinite(const(torch.zeros(3, 2))).all()


def test_sampling_spec_buffer_clip_and_normalisation():
  spec = specs.BoundedTensorSpec((2,), minimum=0.0, maximum=1.0)
  vmin, vmax = np.array([0.2, 0.0], np.float32), np.array([0.8, 1.0], np.float32)
  norm = stats.MinMaxNormalizationLayer(vmin=vmin, vmax=vmax)
  s = get_sampling_spec.get_sampling_spec(spec, vmin, vmax, 0.05, norm)
  # dim 0: [0.2 - 0.03, 0.8 + 0.03] -> +-1.1 ; dim 1: widened range clipped to the env's [0, 1]
  np.testing.assert_allclose(s.minimum, [-1.1, -1.0], atol=1e-6)
  np.testing.assert_allclose(s.maximum, [1.1, 1.0], atol=1e-6)
