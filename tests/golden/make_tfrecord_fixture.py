"""Cuts the TFRecord fixture of tests/test_tfrecord_cpu.py out of the reference's own test
data (run in the build container, where /root/reference exists):

  ibc/test_data/door-human-v0/door-human-v0_0.tfrecord  (tf-agents example_encoding, 1700
  records, 11 episodes) -> records 180 .. 239 (the end of episode 0, LAST at record 199, and
  the start of episode 1, FIRST at record 200) + the shard's .spec file, byte for byte.

The reader is checked against these real tf-agents-written bytes; the expected values in the
test are read straight out of the record bytes at the float payload's marker, not through
the parser under test.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from ibc_b200.ibc.data import tfrecord  # noqa: E402

SRC = '/root/reference/ibc/test_data/door-human-v0/door-human-v0_0.tfrecord'

if __name__ == '__main__':
  records = list(tfrecord.read_records(SRC))
  assert len(records) == 1700
  dst = os.path.join(HERE, 'door_human_180_240.tfrecord')
  tfrecord.write_records(dst, records[180:240])
  shutil.copyfile(SRC + '.spec', dst + '.spec')
  for name in ('PUSH/oracle_push_0.tfrecord.spec',):
    shutil.copyfile('/root/reference/ibc/test_data/' + name,
                    os.path.join(HERE, 'oracle_push_0.tfrecord.spec'))
  print('wrote', dst, os.path.getsize(dst), 'bytes')
