"""N-dimensional particle task: "go to the first goal, then to the second".

Follows environments/particle/particle.py:34-260 of the reference (dynamics
:175-191, reward / success :197-227, reset :147-165, defaults :84-96) without
gym / tf-agents: a plain class with `reset() -> obs`, `step(a) -> (obs, reward,
done, info)`.  Observation = dict(pos_agent, vel_agent, pos_first_goal,
pos_second_goal), each [n_dim] float32; action = position set-point [n_dim]
in [0, 1].  A PD controller integrates the agent for `repeat_actions` internal
steps of `dt` per environment step.

This is the one task of the reference that can be evaluated closed-loop
offline: everything here runs on the host in NumPy, the
policy under test runs on the GPU.
"""
from __future__ import annotations

import collections
import copy

import numpy as np

from ibc_b200 import gin_lite as gin


@gin.configurable
class ParticleEnv:
  """particle.py:34-227."""

  OBS_KEYS = ('pos_agent', 'vel_agent', 'pos_first_goal', 'pos_second_goal')

  def __init__(self, n_steps=50, n_dim=2, hide_velocity=False, seed=None, dt=0.005,
               repeat_actions=10, k_p=10., k_v=5., goal_distance=0.05):
    self.reset_counter = 0
    self.n_steps = n_steps
    self.goal_distance = goal_distance
    self.n_dim = n_dim
    self.hide_velocity = hide_velocity
    self._rng = np.random.RandomState(seed=seed)
    self.dt = dt
    self.repeat_actions = repeat_actions
    assert int(1 / self.dt) % self.repeat_actions == 0      # particle.py:110
    self.k_p = k_p
    self.k_v = k_v
    self.action_low = np.zeros(n_dim, np.float32)
    self.action_high = np.ones(n_dim, np.float32)
    self.reset()

  # -- specs (shapes only; used to build the network and the sampling spec) ---------
  def observation_keys(self):
    return tuple(k for k in self.OBS_KEYS if not (self.hide_velocity and k == 'vel_agent'))

  def seed(self, seed=None):
    self._rng = np.random.RandomState(seed=seed)

  def reset(self):
    self.reset_counter += 1
    self.steps = 0
    self.obs_log = []
    self.act_log = []
    self.new_actions = []
    obs = collections.OrderedDict()
    obs['pos_agent'] = self._rng.rand(self.n_dim).astype(np.float32)
    obs['vel_agent'] = np.zeros(self.n_dim, np.float32)
    obs['pos_first_goal'] = self._rng.rand(self.n_dim).astype(np.float32)
    obs['pos_second_goal'] = self._rng.rand(self.n_dim).astype(np.float32)
    self.obs_log.append(obs)
    self.min_dist_to_first_goal = np.inf
    self.min_dist_to_second_goal = np.inf
    return self._get_state()

  def _get_state(self):
    obs = copy.deepcopy(self.obs_log[-1])
    if self.hide_velocity:
      del obs['vel_agent']
    return obs

  def _internal_step(self, action, new_action):
    if new_action:
      self.new_actions.append(len(self.act_log))
    self.act_log.append({'pos_setpoint': action})
    obs = self.obs_log[-1]
    # u = k_p (x_des - x) - k_v xdot   (desired velocity is zero)
    u_agent = self.k_p * (action - obs['pos_agent']) - self.k_v * obs['vel_agent']
    new_pos = obs['pos_agent'] + obs['vel_agent'] * self.dt
    new_vel = obs['vel_agent'] + u_agent * self.dt
    # a fresh dict per internal step; the goal arrays are never modified in place
    obs = collections.OrderedDict(obs)
    obs['pos_agent'] = new_pos
    obs['vel_agent'] = new_vel
    self.obs_log.append(obs)

  def dist(self, goal):
    return np.linalg.norm(self.obs_log[-1]['pos_agent'] - goal)

  def _get_reward(self, done):
    """1.0 if the agent has hit both goals by the end of the episode."""
    self.min_dist_to_first_goal = min(self.dist(self.obs_log[0]['pos_first_goal']),
                                      self.min_dist_to_first_goal)
    self.min_dist_to_second_goal = min(self.dist(self.obs_log[0]['pos_second_goal']),
                                       self.min_dist_to_second_goal)
    hit = (self.min_dist_to_first_goal < self.goal_distance and
           self.min_dist_to_second_goal < self.goal_distance)
    return 1.0 if (hit and done) else 0.0

  @property
  def succeeded(self):
    thresh = self.goal_distance
    hit_first = self.min_dist_to_first_goal < thresh
    hit_second = self.min_dist_to_second_goal < thresh
    still_at_second = self.dist(self.obs_log[0]['pos_second_goal']) < thresh
    return bool(hit_first and hit_second and still_at_second)

  def step(self, action):
    action = np.asarray(action, np.float32).reshape(self.n_dim)
    self.steps += 1
    self._internal_step(action, new_action=True)
    for _ in range(self.repeat_actions - 1):
      self._internal_step(action, new_action=False)
    state = self._get_state()
    done = self.steps >= self.n_steps
    reward = self._get_reward(done)
    return state, reward, done, {}


class HistoryWrapper:
  """tf_agents HistoryWrapper(history_length, tile_first_step_obs=True) as used by
  ibc/eval/eval_env.py:73-85: observations become [history_length, ...] stacks of
  the most recent steps, the first observation tiled to fill the history."""

  def __init__(self, env, history_length=2):
    self.env = env
    self.history_length = history_length
    self._frames = None

  def __getattr__(self, name):
    return getattr(self.env, name)

  def _stacked(self):
    keys = self._frames[0].keys()
    return collections.OrderedDict(
        (k, np.stack([f[k] for f in self._frames], axis=0)) for k in keys)

  def reset(self):
    obs = self.env.reset()
    self._frames = collections.deque([obs] * self.history_length, maxlen=self.history_length)
    return self._stacked()

  def step(self, action):
    obs, reward, done, info = self.env.step(action)
    self._frames.append(obs)
    return self._stacked(), reward, done, info
