"""Train-and-evaluate loop for the EBM configs (ibc/train_eval.py:86-309 of the reference)
for the task that runs offline, PARTICLE, and -- training only -- for any state-observation
task whose demonstrations are given as TFRecord shards (`--dataset_path`, read without
TensorFlow by ibc_b200/ibc/data/tfrecord.py; e.g. the D4RL door-human shard format).

Same sequence as the reference: demonstration data -> normalisers (get_normalizers) ->
sampling spec (get_sampling_spec) -> MLPEBM (gin-configured) -> ImplicitBCAgent (get_agent)
-> `num_iterations` agent.train steps, closed-loop evaluation of the greedy policy every
`eval_interval` steps on a HistoryWrapper'd ParticleEnv with the particle metrics.  The
tf-agents Learner / Actor / checkpointer / summary plumbing is out of scope; results are
returned as a dict and printed as JSON lines.

  python -m ibc_b200.ibc.train_eval --gin_file configs/particle_mlp_ebm_langevin.gin \
      --gin_bindings "train_eval.num_iterations=2000"
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np
import torch

from ibc_b200 import gin_lite as gin
from ibc_b200.environments.particle import HistoryWrapper, ParticleEnv
from ibc_b200.environments.particle.particle_metrics import ParticleMetrics
from ibc_b200.ibc import specs
from ibc_b200.ibc.data import in_memory
from ibc_b200.ibc.train import get_agent as agent_module
from ibc_b200.ibc.train import get_normalizers as normalizers_module
from ibc_b200.ibc.train import get_sampling_spec as sampling_spec_module
from ibc_b200.networks.mlp_ebm import MLPEBM


def evaluation_step(eval_episodes, eval_env, policy, device):
  """train_eval.py:321-340: run `eval_episodes` episodes with the greedy policy."""
  metrics = ParticleMetrics()
  for _ in range(eval_episodes):
    obs = eval_env.reset()
    step_type, done = 0, False
    while not done:
      batched = {k: torch.as_tensor(v)[None].to(device) for k, v in obs.items()}
      ts = specs.TimeStep(step_type=step_type, reward=0.0, discount=1.0, observation=batched)
      action = policy.action(ts).action.reshape(-1).cpu().numpy()
      obs, _, done, _ = eval_env.step(action)
      step_type = 1
    metrics.record_episode_end(eval_env.env)
  return metrics.result()


@gin.configurable
def train_eval(task='PARTICLE', root_dir=None, dataset_path=None, loss_type='ebm',
               network='MLPEBM', batch_size=512, num_iterations=20000, learning_rate=1e-3,
               decay_steps=100, replay_capacity=100000, eval_interval=1000, eval_episodes=20,
               fused_train_steps=100, sequence_length=2, uniform_boundary_buffer=0.05,
               goal_tolerance=0.02, seed=0, dataset_eval_fraction=0.0, image_obs=False,
               num_demo_episodes=2000, precision=None, log=print):
  del root_dir, replay_capacity, fused_train_steps, goal_tolerance
  del dataset_eval_fraction
  if task != 'PARTICLE' and not dataset_path:
    raise NotImplementedError('only the PARTICLE task can be built and evaluated offline '
                              '(block pushing needs pybullet, D4RL needs mujoco); other tasks '
                              'train from TFRecord shards given as dataset_path')
  if network != 'MLPEBM' or image_obs:
    raise ValueError('train_eval builds the MLPEBM on state observations')
  torch.manual_seed(seed)
  np.random.seed(seed)
  device = torch.device('cuda', torch.cuda.current_device())

  # demonstrations and their sequence windows: TFRecord shards written by the reference's
  # data/policy_eval.py (train_eval.py:143-155, data/dataset.py) when dataset_path is given,
  # else the scripted particle oracle rolled out here (collect_data.sh)
  evaluate = task == 'PARTICLE'
  probe_env = ParticleEnv(seed=seed) if evaluate else None
  if dataset_path:
    import glob
    from ibc_b200.ibc.data import tfrecord
    shards = sorted(p for pat in dataset_path.split(',') for p in glob.glob(pat)
                    if not p.endswith('.spec'))
    if not shards:
      raise ValueError('no TFRecord shards match %r' % dataset_path)
    episodes = tfrecord.episodes_from_shards(shards)
    n_dim = episodes[0][1].shape[1]
    flat_spec = tfrecord.flatten_spec(tfrecord.parse_spec_file(shards[0] + '.spec'))
    act_leaf = flat_spec.get('action')
    act_low = act_leaf.minimum if act_leaf is not None and act_leaf.minimum is not None else None
    act_high = act_leaf.maximum if act_leaf is not None and act_leaf.maximum is not None else None
  else:
    n_dim = probe_env.n_dim
    episodes = in_memory.collect_particle_episodes(num_demo_episodes, n_dim=n_dim, seed=seed)
    act_low = act_high = None
  observations, actions = in_memory.windows_from_episodes(episodes, sequence_length)
  if evaluate:
    act_low, act_high = probe_env.action_low, probe_env.action_high
  else:     # unbounded spec (D4RL shards): the data's own range, as the sampling spec clips to it
    act_low = actions.min(axis=0) if act_low is None else np.broadcast_to(act_low, (n_dim,))
    act_high = actions.max(axis=0) if act_high is None else np.broadcast_to(act_high, (n_dim,))

  # normalisers + sampling spec (train_eval.py:157-184)
  norm_info, norm_train_data_fn = normalizers_module.get_normalizers(
      observations, actions, batch_size, env_name='Particle-v0' if evaluate else task,
      # shards always arrive as a dict here (an un-nested 'observation' tensor becomes a
      # one-key dict), so their normaliser is always the nested form
      **({'nested_obs': True} if dataset_path else {}))
  action_tensor_spec = specs.BoundedTensorSpec((n_dim,), minimum=act_low, maximum=act_high,
                                               name='action')
  action_sampling_spec = sampling_spec_module.get_sampling_spec(
      action_tensor_spec, min_actions=norm_info.min_actions, max_actions=norm_info.max_actions,
      uniform_boundary_buffer=uniform_boundary_buffer, act_norm_layer=norm_info.act_norm_layer)
  log(json.dumps({'action_sampling_spec': {'minimum': action_sampling_spec.minimum.tolist(),
                                           'maximum': action_sampling_spec.maximum.tolist()}}))
  obs_t = {k: torch.as_tensor(v) for k, v in observations.items()}
  (norm_obs, norm_act), _ = norm_train_data_fn((obs_t, torch.as_tensor(actions)))
  data = in_memory.ReplayBatches(norm_obs, norm_act, batch_size, device, seed=seed)

  # network + agent (train_eval.py:190-217)
  obs_tensor_spec = {k: specs.TensorSpec((sequence_length, v.shape[-1]), name=k)
                     for k, v in observations.items()}
  time_step_tensor_spec = specs.TimeStep(step_type=(), reward=(), discount=(),
                                         observation=obs_tensor_spec)
  cloning_network = MLPEBM((obs_tensor_spec, action_tensor_spec), specs.TensorSpec((1,)),
                           seed=seed)
  if precision is not None:          # 'fp32' forces the exact FMA kernels (A/B runs)
    from ibc_b200 import _lib
    cloning_network.engine.set_precision(
        {'fp32': _lib.PRECISION_FP32, 'tf32': _lib.PRECISION_TF32}[precision])
  agent = agent_module.get_agent(
      loss_type, time_step_tensor_spec, action_tensor_spec, action_sampling_spec,
      norm_info.obs_norm_layer, norm_info.act_norm_layer, norm_info.act_denorm_layer,
      learning_rate, False, cloning_network, 0, decay_steps)

  eval_env = (HistoryWrapper(ParticleEnv(n_dim=n_dim, seed=seed + 1000), sequence_length)
              if evaluate else None)
  history = []
  t0 = time.perf_counter()
  step = 0
  while step < num_iterations:
    for _ in range(min(eval_interval, num_iterations - step)):
      info = agent.train(data.next())
      step += 1
    torch.cuda.synchronize()
    row = {'step': step, 'loss': float(info.loss), 'train_seconds': time.perf_counter() - t0}
    row.update({k: float(v) for k, v in agent.last_losses_dict.items()})
    if evaluate:
      row.update(evaluation_step(eval_episodes, eval_env, agent.policy, device))
    history.append(row)
    log(json.dumps(row))
  return {'history': history, 'final': history[-1] if history else None,
          'num_windows': len(data), 'n_dim': n_dim}


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--gin_file', action='append', default=[])
  ap.add_argument('--gin_bindings', action='append', default=[])
  ap.add_argument('--task', default='PARTICLE')
  ap.add_argument('--dataset_path', default=None,
                  help='comma-separated globs of TFRecord shards (with their .spec files)')
  args = ap.parse_args(argv)
  gin.parse_config_files_and_bindings(args.gin_file, args.gin_bindings, skip_unknown=True)
  kwargs = {'task': args.task}
  if args.dataset_path:
    kwargs['dataset_path'] = args.dataset_path
  out = train_eval(**kwargs)
  print(json.dumps({'final': out['final']}))


if __name__ == '__main__':
  main()
