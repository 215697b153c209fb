"""Scripted expert for the particle task (environments/particle/particle_oracles.py:28-85):
set-point = first goal until the agent has been within `goal_threshold` of it for
`wait_at_first_goal` steps, then the second goal.  `multimodal` shuffles the goal
order per episode."""
from __future__ import annotations

import random

import numpy as np


class ParticleOracle:

  def __init__(self, env, wait_at_first_goal=1, multimodal=False, goal_threshold=0.01):
    assert wait_at_first_goal > 0
    assert goal_threshold > 0.
    self._env = env
    self.wait_at_first_goal = wait_at_first_goal
    self.goal_threshold = goal_threshold
    self.multimodal = multimodal
    self.reset()

  def reset(self):
    self.steps_at_first_goal = 0
    self.goal_order = ['pos_first_goal', 'pos_second_goal']
    if self.multimodal:
      random.shuffle(self.goal_order)

  def action(self, obs, is_first=False):
    if is_first:
      self.reset()
    first_key, second_key = self.goal_order
    gt_goals = self._env.obs_log[0]
    dist = np.linalg.norm(obs['pos_agent'] - gt_goals[first_key])
    if dist < self.goal_threshold:
      self.steps_at_first_goal += 1
    if self.steps_at_first_goal < self.wait_at_first_goal:
      return np.copy(gt_goals[first_key])
    return np.copy(gt_goals[second_key])
