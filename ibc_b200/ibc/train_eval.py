"""Train-and-evaluate loop for the EBM configs (ibc/train_eval.py:86-309 of the reference)
for the task that runs offline, PARTICLE, and -- training only -- for any state-observation
task whose demonstrations are given as TFRecord shards (`--dataset_path`, read without
TensorFlow by ibc_b200/ibc/data/tfrecord.py; e.g. the D4RL door-human shard format).

Same sequence as the reference: demonstration data -> normalisers (get_normalizers) ->
sampling spec (get_sampling_spec) -> MLPEBM (gin-configured) -> ImplicitBCAgent (get_agent)
-> `num_iterations` agent.train steps, closed-loop evaluation of the greedy policy every
`eval_interval` steps on a HistoryWrapper'd ParticleEnv with the particle metrics.  Also
from the reference loop: training runs in bursts of `fused_train_steps` (train_eval.py:259-261),
a held-out `dataset_eval_fraction` of the windows gives a validation loss through
`agent.get_eval_loss` every `eval_loss_interval` steps (:262-267, :334-345), the
per-replica batch is `batch_size // num_replicas` under a torch.distributed process group
(:163, one process per GPU, gradients all-reduced inside agent.train), and the training state
-- flat parameters, Adam m / v / iterations, train step, spectral-norm u, data RNG -- is
checkpointed to `root_dir` every `checkpoint_interval` steps and resumed from there
(get_learner.py:60-76: the Learner's checkpointer; tf-agents' SavedModel export and summary
writers are out of scope).  Results are returned as a dict and printed as JSON lines.

  python -m ibc_b200.ibc.train_eval --gin_file configs/particle_mlp_ebm_langevin.gin \
      --gin_bindings "train_eval.num_iterations=2000"
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np
import torch

from ibc_b200 import gin_lite as gin
from ibc_b200 import parallel
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


CHECKPOINT_NAME = 'ibc_b200_checkpoint.pt'


def save_checkpoint(root_dir, agent, data, step):
  """Flat buffers of the training state (get_learner.py:60-76's checkpointer): written to a
  temporary file and renamed, so an interrupted save never leaves a torn checkpoint."""
  net, opt = agent.cloning_network, agent._optimizer
  state = {'step': int(step), 'train_step_counter': int(agent.train_step_counter),
           'params': net.params.detach().cpu(),
           'adam_m': None if opt.m is None else opt.m.detach().cpu(),
           'adam_v': None if opt.v is None else opt.v.detach().cpu(),
           'adam_iterations': int(opt.iterations),
           'sn_u': None if net._sn_u is None else {k: v.detach().cpu() for k, v in net._sn_u.items()},
           'data_rng': data._gen.get_state(), 'torch_rng': torch.get_rng_state()}
  os.makedirs(root_dir, exist_ok=True)
  path = os.path.join(root_dir, CHECKPOINT_NAME)
  torch.save(state, path + '.tmp')
  os.replace(path + '.tmp', path)
  return path


def load_checkpoint(root_dir, agent, data):
  """Restores what save_checkpoint wrote; returns the step to continue from (0 if none)."""
  path = os.path.join(root_dir or '', CHECKPOINT_NAME)
  if not root_dir or not os.path.exists(path):
    return 0
  state = torch.load(path, map_location='cpu', weights_only=False)
  net, opt = agent.cloning_network, agent._optimizer
  if state['params'].numel() != net.params.numel():
    raise ValueError('checkpoint %s holds %d parameters, the network has %d' % (
        path, state['params'].numel(), net.params.numel()))
  if state['sn_u'] is not None and net._sn_u is not None:
    net._sn_u = {k: v.to(net.device) for k, v in state['sn_u'].items()}
  net.set_params(state['params'])
  if state['adam_m'] is not None:
    opt.m = state['adam_m'].to(net.device)
    opt.v = state['adam_v'].to(net.device)
  opt.iterations = state['adam_iterations']
  agent.train_step_counter = state['train_step_counter']
  data._gen.set_state(state['data_rng'])
  torch.set_rng_state(state['torch_rng'])
  return int(state['step'])


def validation_step(eval_data, agent):
  """train_eval.py:334-345: one batch of the held-out windows through agent.get_eval_loss
  (the same _loss with training=False, base_agent.py:64-66)."""
  losses = agent.get_eval_loss(eval_data.next())
  return {'eval_' + k: float(v) for k, v in losses.items()}


@gin.configurable
def train_eval(task='PARTICLE', root_dir=None, dataset_path=None, loss_type='ebm',
               network='MLPEBM', batch_size=512, num_iterations=20000, learning_rate=1e-3,
               decay_steps=100, replay_capacity=100000, eval_interval=1000, eval_episodes=20,
               eval_loss_interval=100, fused_train_steps=100, checkpoint_interval=5000,
               sequence_length=2, uniform_boundary_buffer=0.05,
               goal_tolerance=0.02, seed=0, dataset_eval_fraction=0.0, image_obs=False,
               num_demo_episodes=2000, precision=None, log=print):
  del replay_capacity, goal_tolerance
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
  # held-out windows for the validation loss (get_data.py:41: eval_fraction of the data)
  eval_data = None
  if dataset_eval_fraction > 0.0:
    n_all = norm_act.shape[0]
    perm = torch.randperm(n_all, generator=torch.Generator().manual_seed(seed + 17))
    n_eval = max(1, int(n_all * dataset_eval_fraction))
    ev, tr = perm[:n_eval], perm[n_eval:]
    eval_obs, eval_act = {k: v[ev] for k, v in norm_obs.items()}, norm_act[ev]
    norm_obs, norm_act = {k: v[tr] for k, v in norm_obs.items()}, norm_act[tr]
  # one process per GPU: every replica draws its own per_replica_batch_size windows
  # (train_eval.py:163) from its own stream; the loss is scaled by the global batch and the
  # gradients are all-reduced inside agent.train
  replicas, rank = parallel.num_replicas(), parallel.replica_rank()
  per_replica_batch_size = parallel.per_replica_batch_size(batch_size)
  data = in_memory.ReplayBatches(norm_obs, norm_act, per_replica_batch_size, device,
                                 seed=seed + 7919 * rank)
  if dataset_eval_fraction > 0.0:
    eval_data = in_memory.ReplayBatches(eval_obs, eval_act, per_replica_batch_size, device,
                                        seed=seed + 31 + 7919 * rank)

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
  step = load_checkpoint(root_dir, agent, data)          # resume (get_learner.py:60-76)
  if step:
    log(json.dumps({'resumed_from_step': step}))
  # the interpreter's start-up heap out of the cyclic collector's way: a full collection over it is
  # a pause of tens of milliseconds, which in a data-parallel run every other rank waits for
  import gc
  gc.collect()
  gc.freeze()
  t0 = time.perf_counter()
  info = None
  while step < num_iterations:
    # one burst of at most fused_train_steps (training_step, train_eval.py:312-318), cut at the
    # next evaluation / validation / checkpoint boundary when the intervals are not multiples
    stops = [num_iterations, step + fused_train_steps, (step // eval_interval + 1) * eval_interval]
    if eval_data is not None:
      stops.append((step // eval_loss_interval + 1) * eval_loss_interval)
    if root_dir:
      stops.append((step // checkpoint_interval + 1) * checkpoint_interval)
    for _ in range(min(stops) - step):
      info = agent.train(data.next())
      step += 1
    row = None
    if eval_data is not None and step % eval_loss_interval == 0:
      row = {'step': step}
      row.update(validation_step(eval_data, agent))
    if step % eval_interval == 0 or step >= num_iterations:
      torch.cuda.synchronize()
      row = row or {'step': step}
      row.update({'loss': float(info.loss), 'train_seconds': time.perf_counter() - t0,
                  'replicas': replicas, 'per_replica_batch_size': per_replica_batch_size})
      row.update({k: float(v) for k, v in agent.last_losses_dict.items()})
      if evaluate and rank == 0:
        row.update(evaluation_step(eval_episodes, eval_env, agent.policy, device))
    if row is not None:
      history.append(row)
      if rank == 0:
        log(json.dumps(row))
    # after the evaluation, which draws from the sampler's RNG stream: a resumed run then
    # continues exactly where the uninterrupted one would be
    if root_dir and rank == 0 and (step % checkpoint_interval == 0 or step >= num_iterations):
      save_checkpoint(root_dir, agent, data, step)
  if hasattr(agent, 'flush_checks'):
    agent.flush_checks()
  final = next((r for r in reversed(history) if 'loss' in r), None)
  return {'history': history, 'final': final, 'num_windows': len(data), 'n_dim': n_dim,
          'agent': agent}


def main(argv=None):
  ap = argparse.ArgumentParser()
  ap.add_argument('--gin_file', action='append', default=[])
  ap.add_argument('--gin_bindings', action='append', default=[])
  ap.add_argument('--task', default='PARTICLE')
  ap.add_argument('--dataset_path', default=None,
                  help='comma-separated globs of TFRecord shards (with their .spec files)')
  ap.add_argument('--root_dir', default=None, help='checkpoints are written here and resumed from')
  args = ap.parse_args(argv)
  if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
    import torch.distributed as dist          # launched by torchrun: one process per GPU
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  gin.parse_config_files_and_bindings(args.gin_file, args.gin_bindings, skip_unknown=True)
  kwargs = {'task': args.task}
  if args.dataset_path:
    kwargs['dataset_path'] = args.dataset_path
  if args.root_dir:
    kwargs['root_dir'] = args.root_dir
  out = train_eval(**kwargs)
  if parallel.replica_rank() == 0:
    print(json.dumps({'final': out['final']}))


if __name__ == '__main__':
  main()
