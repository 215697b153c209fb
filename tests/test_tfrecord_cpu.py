"""CPU tests of the TensorFlow-free TFRecord reader (ibc_b200/ibc/data/tfrecord.py) against
real tf-agents-written bytes (tests/golden/door_human_180_240.tfrecord, cut from the
reference's ibc/test_data by tests/golden/make_tfrecord_fixture.py) and against the windowing
rules of data/dataset.py:29-84, 114-196."""
import os

import numpy as np
import pytest

from ibc_b200.ibc.data import tfrecord

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
SHARD = os.path.join(GOLDEN, 'door_human_180_240.tfrecord')


def test_crc32c_known_answers():
  # RFC 3720 B.4 test vectors
  assert tfrecord.crc32c(b'\x00' * 32) == 0x8A9136AA
  assert tfrecord.crc32c(b'\xff' * 32) == 0x62A8AB43
  assert tfrecord.crc32c(bytes(range(32))) == 0x46DD794E
  assert tfrecord.crc32c(b'123456789') == 0xE3069283


def test_spec_file_of_a_d4rl_shard():
  spec = tfrecord.flatten_spec(tfrecord.parse_spec_file(SHARD + '.spec'))
  assert list(spec) == ['step_type', 'observation', 'action', 'next_step_type', 'reward', 'discount']
  assert spec['observation'].shape == (39,) and spec['observation'].dtype == np.float32
  assert spec['action'].shape == (28,) and spec['action'].dtype == np.float32
  assert spec['step_type'].shape == () and spec['step_type'].dtype == np.int32


def test_spec_file_with_dict_observations_and_bounds():
  """Block pushing: dict observation of bounded specs (sorted keys, '/'-joined paths)."""
  spec = tfrecord.parse_spec_file(os.path.join(GOLDEN, 'oracle_push_0.tfrecord.spec'))
  assert list(spec['observation']) == [
      'block_orientation', 'block_translation', 'effector_target_translation',
      'effector_translation', 'target_orientation', 'target_translation']
  flat = tfrecord.flatten_spec(spec)
  assert flat['observation/block_translation'].shape == (2,)
  np.testing.assert_allclose(flat['action'].minimum, -0.1, rtol=1e-6)
  np.testing.assert_allclose(flat['action'].maximum, 0.1, rtol=1e-6)
  np.testing.assert_allclose(flat['observation/effector_translation'].minimum, [0.05, -0.6], rtol=1e-6)
  np.testing.assert_allclose(flat['observation/effector_translation'].maximum, [0.8, 0.6], rtol=1e-6)


def test_records_decode_to_the_bytes_tf_agents_wrote():
  records = list(tfrecord.read_records(SHARD, verify=True))
  assert len(records) == 60
  data = tfrecord.load_shard(SHARD)
  assert data['observation'].shape == (60, 39) and data['action'].shape == (60, 28)
  # expected values straight from the record bytes: the packed float payload follows the
  # feature's key / length markers (156 = 0x9c01 bytes of observation, 112 = 0x70 of action)
  obs_mark = b'\n\x0bobservation\x12\xa2\x01\x12\x9f\x01\n\x9c\x01'
  act_mark = b'\n\x06action\x12t\x12r\np'
  for i, r in enumerate(records):
    o = r.index(obs_mark) + len(obs_mark)
    a = r.index(act_mark) + len(act_mark)
    np.testing.assert_array_equal(data['observation'][i], np.frombuffer(r[o:o + 156], '<f4'))
    np.testing.assert_array_equal(data['action'][i], np.frombuffer(r[a:a + 112], '<f4'))
  # records 180 .. 239 of the shard: episode 0 ends at record 199, episode 1 starts at 200
  st = data['step_type']
  assert st[19] == tfrecord.STEP_LAST and st[20] == tfrecord.STEP_FIRST
  assert (st[:19] == tfrecord.STEP_MID).all() and (st[21:] == tfrecord.STEP_MID).all()
  assert (data['discount'][:19] == 1.0).all() and np.abs(data['action']).max() <= 1.0


def test_corruption_is_detected(tmp_path):
  raw = bytearray(open(SHARD, 'rb').read())
  raw[40] ^= 0x01
  bad = tmp_path / 'bad.tfrecord'
  bad.write_bytes(bytes(raw))
  with pytest.raises(ValueError):
    list(tfrecord.read_records(str(bad)))
  assert len(list(tfrecord.read_records(str(bad), verify=False))) == 60
  with pytest.raises(ValueError):
    trunc = tmp_path / 'trunc.tfrecord'
    trunc.write_bytes(bytes(raw[:1000]))
    list(tfrecord.read_records(str(trunc), verify=False))


def test_example_round_trip():
  feats = {'f': np.array([1.5, -2.25, 0.0], np.float32), 'i': np.array([0, -1, 2 ** 40], np.int64),
           'b': [np.arange(3, dtype='<i4').tobytes()]}
  back = tfrecord.parse_example(tfrecord.encode_example(feats))
  np.testing.assert_array_equal(back['f'], feats['f'])
  np.testing.assert_array_equal(back['i'], feats['i'])
  assert back['b'] == feats['b']


def test_filter_episodes_indices():
  """data/dataset.py:54-65: [FIRST, MID, LAST, FIRST, MID, MID] -> [3, 3, 3, 3, 4, 5]."""
  F, M, L = tfrecord.STEP_FIRST, tfrecord.STEP_MID, tfrecord.STEP_LAST
  np.testing.assert_array_equal(
      tfrecord.filter_episodes_indices(np.array([F, M, L, F, M, M])), [3, 3, 3, 3, 4, 5])
  np.testing.assert_array_equal(tfrecord.filter_episodes_indices(np.array([M, M, L])), [0, 1, 2])
  np.testing.assert_array_equal(tfrecord.filter_episodes_indices(np.array([F, M, M])), [0, 1, 2])
  np.testing.assert_array_equal(tfrecord.filter_episodes_indices(np.array([M, F])), [1, 1])


def test_sequence_windows_follow_the_reference():
  """'a sequence of [0, 1, 2] with seq_len = 2 produces [0, 1], [1, 2], [0, 0]' per shard
  (data/dataset.py:122-125): windows of adjacent records, wrapping around the shard, the
  new episode's first frame replicated over the frames of the previous episode; the action
  is the last step's (get_data.py:57-59)."""
  data = tfrecord.load_shard(SHARD)
  obs, act, st = tfrecord.load_tfrecord_dataset_sequence([SHARD], seq_len=3)
  assert obs['observation'].shape == (60, 3, 39) and act.shape == (60, 28)
  np.testing.assert_array_equal(obs['observation'][0], data['observation'][0:3])
  np.testing.assert_array_equal(act[0], data['action'][2])
  # window starting at record 18: [MID, LAST, FIRST] -> frame 20 three times
  np.testing.assert_array_equal(st[18], [0, 0, 0])
  np.testing.assert_array_equal(obs['observation'][18], np.stack([data['observation'][20]] * 3))
  np.testing.assert_array_equal(act[18], data['action'][20])
  # window starting at record 19: [LAST, FIRST, MID] -> [20, 20, 21]
  np.testing.assert_array_equal(obs['observation'][19],
                                data['observation'][[20, 20, 21]])
  # the last window wraps to the start of the shard, which is not a FIRST frame: kept as is
  np.testing.assert_array_equal(obs['observation'][59], data['observation'][[59, 0, 1]])


def test_episodes_feed_the_in_memory_windows():
  from ibc_b200.ibc.data import in_memory
  episodes = tfrecord.episodes_from_shards([SHARD])
  assert [e[1].shape[0] for e in episodes] == [20, 40]
  observations, actions = in_memory.windows_from_episodes(episodes, sequence_length=2)
  assert observations['observation'].shape == (60, 2, 39) and actions.shape == (60, 28)
  # first window of the second episode: its first frame twice
  np.testing.assert_array_equal(observations['observation'][20][0], observations['observation'][20][1])
