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
