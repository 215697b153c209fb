"""GPU test of the train_eval loop on the particle task (ibc/train_eval.py, SURVEY.md
8f-2): a short run of the reference's particle Langevin config wires data ->
normalisers -> sampling spec -> MLPEBM -> ImplicitBCAgent -> closed-loop evaluation."""
import os

import pytest

pytestmark = pytest.mark.gpu

_GIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'configs',
                    'particle_mlp_ebm_langevin.gin')


def test_particle_train_eval_short_run():
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import train_eval
  gin.clear_config()
  gin.parse_config_files_and_bindings(
      [_GIN], ['train_eval.num_iterations = 120', 'train_eval.eval_interval = 60',
               'train_eval.eval_episodes = 2', 'train_eval.num_demo_episodes = 100',
               'train_eval.batch_size = 128', 'langevin_actions_given_obs.num_iterations = 20'])
  try:
    out = train_eval.train_eval(log=lambda s: None)
  finally:
    gin.clear_config()
  hist = out['history']
  assert [r['step'] for r in hist] == [60, 120]
  assert out['num_windows'] == 100 * 50 and out['n_dim'] == 2
  for r in hist:
    for key in ('loss', 'ebm_total_loss', 'grad_loss', 'mse_counter_examples',
                'AverageSuccessMetric', 'AverageFirstGoalDistance',
                'AverageSecondGoalDistance', 'AverageFinalSecondGoalDistance'):
      assert key in r and r[key] == r[key], key        # present and not NaN
    assert 0.0 <= r['AverageSuccessMetric'] <= 1.0
  # InfoNCE over 9 candidates starts near log 9 and falls as the EBM fits the oracle
  assert hist[-1]['ebm_total_loss'] < 2.1972


def test_train_from_tfrecord_shards_d4rl_shape():
  """a D4RL-shaped run (door-human: 39-d observation, 28-d action, the
  d4rl Langevin gin's 512-wide 8-deep MLPEBM) trained from a tf-agents TFRecord shard read
  without TensorFlow (tests/golden/door_human_180_240.tfrecord, real reference bytes)."""
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import train_eval
  shard = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                       'door_human_180_240.tfrecord')
  gin.clear_config()
  gin.parse_config_files_and_bindings(
      [], ['train_eval.task = "door-human-v0"', 'train_eval.dataset_path = "%s"' % shard,
           'train_eval.sequence_length = 1', 'train_eval.num_iterations = 40',
           'train_eval.eval_interval = 20', 'train_eval.batch_size = 32',
           'MLPEBM.width = 512', 'MLPEBM.depth = 8', 'MLPEBM.layers = "ResNetPreActivation"',
           'MLPEBM.rate = 0.0', 'ImplicitBCAgent.num_counter_examples = 8',
           'ImplicitBCAgent.fraction_langevin_samples = 1.0',
           'ImplicitBCAgent.add_grad_penalty = True',
           'langevin_actions_given_obs.num_iterations = 10'])
  try:
    out = train_eval.train_eval(log=lambda s: None)
  finally:
    gin.clear_config()
  assert out['num_windows'] == 60 and out['n_dim'] == 28
  hist = out['history']
  assert [r['step'] for r in hist] == [20, 40]
  for r in hist:
    assert r['loss'] == r['loss'] and 'AverageSuccessMetric' not in r
  assert hist[-1]['ebm_total_loss'] < 2.1972


def test_checkpoint_resume_and_validation_loss(tmp_path):
  """checkpoint / resume of the flat training state (get_learner.py:60-76) and
  the validation loss on held-out windows (train_eval.py:262-267).  A run to step 60 that was
  checkpointed at step 30 and resumed must end with EXACTLY the parameters of an uninterrupted
  run -- parameters, Adam moments and iteration count (hence the learning-rate schedule), data
  and sampler RNG streams all restored.  The fp32 path is used (its kernels sum without
  atomics), uniform counter-examples."""
  import torch
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import train_eval

  def run(root, iters):
    gin.clear_config()
    gin.parse_config_files_and_bindings(
        [], ['train_eval.num_iterations = %d' % iters, 'train_eval.eval_interval = 30',
             'train_eval.eval_episodes = 1', 'train_eval.num_demo_episodes = 40',
             'train_eval.batch_size = 64', 'train_eval.fused_train_steps = 10',
             'train_eval.checkpoint_interval = 30', 'train_eval.eval_loss_interval = 20',
             'train_eval.dataset_eval_fraction = 0.1', 'train_eval.precision = "fp32"',
             'train_eval.root_dir = "%s"' % root, 'get_normalizers.nested_obs = True',
             'MLPEBM.width = 64', 'MLPEBM.depth = 2', 'MLPEBM.layers = "ResNetPreActivation"',
             'MLPEBM.rate = 0.0', 'ImplicitBCAgent.num_counter_examples = 16',
             'ImplicitBCAgent.fraction_langevin_samples = 0.0',
             'ImplicitBCAgent.add_grad_penalty = False',
             'IbcPolicy.num_action_samples = 256', 'IbcPolicy.use_dfo = True',
             'IbcPolicy.use_langevin = False'])
    try:
      return train_eval.train_eval(log=lambda s: None)
    finally:
      gin.clear_config()

  full = run(str(tmp_path / 'a'), 60)
  steps = [r['step'] for r in full['history']]
  assert steps == [20, 30, 40, 60], steps
  assert all(('eval_ebm_total_loss' in r) == (r['step'] % 20 == 0) for r in full['history'])
  assert os.path.exists(str(tmp_path / 'a' / train_eval.CHECKPOINT_NAME))
  half = run(str(tmp_path / 'b'), 30)
  assert half['agent'].train_step_counter == 30
  resumed = run(str(tmp_path / 'b'), 60)
  assert [r['step'] for r in resumed['history']] == [40, 60]
  assert resumed['agent'].train_step_counter == 60
  assert resumed['agent']._optimizer.iterations == 60
  a, b = resumed['agent'].cloning_network.params, full['agent'].cloning_network.params
  # Identical up to the summation order of the atomically accumulated gradient entries, which
  # differs from run to run (alone in a process the two runs happened to agree to 1e-8; behind
  # the rest of the suite the median difference was 3e-7).  Adam turns that rounding noise into
  # steps of +-lr on entries whose gradient is pure noise: Dense(1)'s bias, the last entry
  # (its gradient is the sum of all dE = 0), and the top block's b2 = (sum dE) x we.  A resume
  # that lost the moments, the step count or an RNG stream trains on for 30 steps of lr = 1e-3
  # from a different state: differences of 1e-3 .. 1e-2 on most entries.
  d = (a - b).abs()
  # (the noise-driven entries -- be and the 64 of b2 -- are 2.4 % of this net's parameters: the
  # fraction above 1e-4 comes out at either ~0 or 0.0243 depending on the run)
  assert float(d.median()) <= 1e-5 and float((d > 1e-4).float().mean()) <= 0.05, (
      float(d.median()), float((d > 1e-4).float().mean()))
  assert float(d[:-1].max()) <= 2e-3 and float(d[-1]) <= 60 * 1e-3, float(d.max())
