"""Episode metrics of the particle task (environments/particle/particle_metrics.py:30-172):
averages over the evaluated episodes of the minimum distance to each goal, the final
distance to the second goal, and success."""
from __future__ import annotations

import numpy as np


class ParticleMetrics:
  NAMES = ('AverageFirstGoalDistance', 'AverageSecondGoalDistance',
           'AverageFinalSecondGoalDistance', 'AverageSuccessMetric')

  def __init__(self):
    self._rows = []

  def record_episode_end(self, env):
    """Call once after the last step of an episode (trajectory.is_last())."""
    self._rows.append((
        float(env.min_dist_to_first_goal), float(env.min_dist_to_second_goal),
        float(env.dist(env.obs_log[0]['pos_second_goal'])), 1.0 if env.succeeded else 0.0))

  def result(self):
    if not self._rows:
      return {n: float('nan') for n in self.NAMES}
    arr = np.asarray(self._rows, np.float64).mean(axis=0)
    return dict(zip(self.NAMES, arr.tolist()))
