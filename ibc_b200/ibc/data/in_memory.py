"""In-memory demonstration data for the offline-runnable task.

The reference collects oracle episodes into TFRecords with `data/policy_eval.py`
(`ibc/configs/particle/collect_data.sh`: 200 episodes x 10 replicas of the scripted
particle oracle) and reads them back as windows of `sequence_length` consecutive steps
(`ibc/train/get_data.py:28-70`: observation window [T, ...], action of the last step).
TFRecords need TensorFlow; the same episodes are generated here directly and the same
windows are cut from them.
"""
from __future__ import annotations

import numpy as np
import torch

from ibc_b200.environments.particle import ParticleEnv, ParticleOracle


def collect_particle_episodes(num_episodes, n_dim=2, seed=0, **oracle_kwargs):
  """Roll the scripted oracle out; returns a list of (obs dict of [L+1, d], actions [L, d])."""
  env = ParticleEnv(n_dim=n_dim, seed=seed)
  oracle = ParticleOracle(env, **oracle_kwargs)
  episodes = []
  for _ in range(num_episodes):
    obs = env.reset()
    oracle.reset()
    obs_seq, act_seq = [obs], []
    done = False
    while not done:
      act = oracle.action(obs).astype(np.float32)
      obs, _, done, _ = env.step(act)
      obs_seq.append(obs)
      act_seq.append(act)
    keys = obs_seq[0].keys()
    episodes.append(({k: np.stack([o[k] for o in obs_seq]) for k in keys}, np.stack(act_seq)))
  return episodes


def windows_from_episodes(episodes, sequence_length=2):
  """One window per step t of every episode: the `sequence_length` most recent
  observations, the episode's first frame replicated where the window would reach
  before it (data/dataset.py:29-84, 114-125 of the reference: "a sequence of [0, 1, 2]
  with seq_len = 2 produces [0, 0], [0, 1], [1, 2]" -- the same padding the evaluation
  environment's HistoryWrapper applies), and the action taken at step t
  (get_data.py:57-59).  Returns observations [S, T, d] per key and actions [S, A]."""
  obs_out, act_out = None, []
  for obs, act in episodes:
    steps = act.shape[0]
    for t in range(steps):
      idx = np.maximum(np.arange(t - sequence_length + 1, t + 1), 0)
      if obs_out is None:
        obs_out = {k: [] for k in obs}
      for k in obs:
        obs_out[k].append(obs[k][idx])
      act_out.append(act[t])
  observations = {k: np.stack(v).astype(np.float32) for k, v in obs_out.items()}
  return observations, np.stack(act_out).astype(np.float32)


class ReplayBatches:
  """Uniform-with-replacement batches from device-resident normalised windows (the
  reference's shuffled, repeated tf.data stream with `replay_capacity`)."""

  def __init__(self, observations, actions, batch_size, device, seed=0):
    self.obs = {k: torch.as_tensor(v).to(device) for k, v in observations.items()}
    self.act = torch.as_tensor(actions).to(device)
    self.batch_size = batch_size
    self._gen = torch.Generator(device='cpu').manual_seed(seed)
    self._device = device

  def __len__(self):
    return self.act.shape[0]

  def next(self):
    idx = torch.randint(0, len(self), (self.batch_size,), generator=self._gen).to(self._device)
    return ({k: v[idx] for k, v in self.obs.items()}, self.act[idx]), ()
