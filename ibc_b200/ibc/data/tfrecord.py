"""TFRecord reader for tf-agents trajectory datasets without TensorFlow.

The reference stores demonstrations as TFRecord shards of `tf.Example` protos written by
tf-agents' `example_encoding_dataset` next to a `.spec` file (a one-record TFRecord holding
a `StructuredValue` proto of the `Trajectory` spec), and reads them back as windows of
`seq_len` adjacent records (`data/dataset.py:114-196`, `filter_episodes` `:29-84`), keeping
the observation window and the action of the last step (`ibc/train/get_data.py:57-59`).
This module restates exactly that on NumPy:

  * record framing: `uint64 length | uint32 masked crc32c(length) | data | uint32 masked
    crc32c(data)` (TensorFlow's `tensorflow/core/lib/io/record_writer.cc`);
  * `tf.Example`: `features { feature { key, Feature{bytes_list=1 | float_list=2 |
    int64_list=3} } }`; tf-agents `example_encoding.py` writes float32 tensors as
    `float_list`, int64 as `int64_list` and every other dtype as the tensor's raw
    little-endian bytes in a one-element `bytes_list`;
  * the spec: `StructuredValue` (`tensorflow/core/protobuf/struct.proto`) with
    `named_tuple_value` (54), `dict_value` (53), `tuple_value` (52), `list_value` (51),
    `tensor_spec_value` (33), `bounded_tensor_spec_value` (35).

Feature keys are the '/'-joined paths of the spec nest (`observation/pos_agent`, `action`).
JPEG / PNG compressed image features (`compress_image=True` writers) are not decoded here.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Sequence, Tuple

import numpy as np

# ---- crc32c (Castagnoli), masked as TensorFlow does -----------------------------------------
_CRC_TABLE = None


def _crc_table():
  global _CRC_TABLE
  if _CRC_TABLE is None:
    poly = 0x82F63B78
    table = []
    for i in range(256):
      c = i
      for _ in range(8):
        c = (c >> 1) ^ poly if c & 1 else c >> 1
      table.append(c)
    _CRC_TABLE = table
  return _CRC_TABLE


def crc32c(data: bytes) -> int:
  table = _crc_table()
  c = 0xFFFFFFFF
  for b in data:
    c = table[(c ^ b) & 0xFF] ^ (c >> 8)
  return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
  c = crc32c(data)
  return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def read_records(path: str, verify: bool = True) -> Iterator[bytes]:
  """Yields the payload of every record of a TFRecord file; `verify` checks both CRCs."""
  with open(path, 'rb') as f:
    buf = f.read()
  pos, n = 0, len(buf)
  while pos < n:
    if pos + 12 > n:
      raise ValueError('%s: truncated record header at byte %d' % (path, pos))
    (length,) = struct.unpack_from('<Q', buf, pos)
    (len_crc,) = struct.unpack_from('<I', buf, pos + 8)
    if verify and masked_crc32c(buf[pos:pos + 8]) != len_crc:
      raise ValueError('%s: corrupt length at byte %d' % (path, pos))
    start, end = pos + 12, pos + 12 + length
    if end + 4 > n:
      raise ValueError('%s: truncated record at byte %d' % (path, pos))
    data = buf[start:end]
    (data_crc,) = struct.unpack_from('<I', buf, end)
    if verify and masked_crc32c(data) != data_crc:
      raise ValueError('%s: corrupt record at byte %d' % (path, pos))
    yield data
    pos = end + 4


def write_records(path: str, records: Sequence[bytes]) -> None:
  """Writes TFRecord framing (used by the tests' own fixtures)."""
  with open(path, 'wb') as f:
    for data in records:
      head = struct.pack('<Q', len(data))
      f.write(head + struct.pack('<I', masked_crc32c(head)) + data +
              struct.pack('<I', masked_crc32c(data)))


# ---- minimal protobuf wire format ------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
  result, shift = 0, 0
  while True:
    b = buf[pos]
    pos += 1
    result |= (b & 0x7F) << shift
    if not b & 0x80:
      return result, pos
    shift += 7


def _fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
  """(field number, wire type, value) of a serialized message; value is an int for
  varint / fixed fields and bytes for length-delimited ones."""
  pos, n = 0, len(buf)
  while pos < n:
    key, pos = _varint(buf, pos)
    field, wire = key >> 3, key & 7
    if wire == 0:
      val, pos = _varint(buf, pos)
    elif wire == 1:
      val = struct.unpack_from('<Q', buf, pos)[0]
      pos += 8
    elif wire == 2:
      length, pos = _varint(buf, pos)
      val = buf[pos:pos + length]
      pos += length
    elif wire == 5:
      val = struct.unpack_from('<I', buf, pos)[0]
      pos += 4
    else:
      raise ValueError('unsupported protobuf wire type %d' % wire)
    yield field, wire, val


def _put_varint(n: int) -> bytes:
  out = bytearray()
  while True:
    b = n & 0x7F
    n >>= 7
    if n:
      out.append(b | 0x80)
    else:
      out.append(b)
      return bytes(out)


def _put_field(field: int, payload: bytes) -> bytes:
  return _put_varint((field << 3) | 2) + _put_varint(len(payload)) + payload


# ---- tf.Example --------------------------------------------------------------------------------
def parse_example(record: bytes) -> Dict[str, object]:
  """tf.Example -> {key: float32 array | int64 array | list of bytes}."""
  out: Dict[str, object] = {}
  for f, _, features in _fields(record):
    if f != 1:
      continue
    for g, _, entry in _fields(features):        # map<string, Feature> entries
      if g != 1:
        continue
      key, feature = None, b''
      for h, _, v in _fields(entry):
        if h == 1:
          key = v.decode('utf-8')
        elif h == 2:
          feature = v
      value: object = []
      for kind, _, payload in _fields(feature):
        if kind == 2:      # FloatList: packed floats (or repeated fixed32)
          floats: List[np.ndarray] = []
          for q, wire, v in _fields(payload):
            if q != 1:
              continue
            floats.append(np.frombuffer(v, dtype='<f4') if wire == 2
                          else np.array([struct.unpack('<f', struct.pack('<I', v))[0]], 'f4'))
          value = np.concatenate(floats) if floats else np.zeros(0, np.float32)
        elif kind == 3:    # Int64List: packed varints
          ints: List[int] = []
          for q, wire, v in _fields(payload):
            if q != 1:
              continue
            if wire == 2:
              p = 0
              while p < len(v):
                x, p = _varint(v, p)
                ints.append(x - (1 << 64) if x >= (1 << 63) else x)
            else:
              ints.append(v - (1 << 64) if v >= (1 << 63) else v)
          value = np.array(ints, dtype=np.int64)
        elif kind == 1:    # BytesList
          value = [v for q, _, v in _fields(payload) if q == 1]
      if key is not None:
        out[key] = value
  return out


def encode_example(features: Dict[str, object]) -> bytes:
  """Inverse of parse_example for float32 / int64 arrays and lists of bytes."""
  entries = b''
  for key, value in features.items():
    if isinstance(value, (list, tuple)):
      feature = _put_field(1, b''.join(_put_field(1, bytes(v)) for v in value))
    else:
      arr = np.asarray(value)
      if arr.dtype.kind == 'f':
        feature = _put_field(2, _put_field(1, arr.astype('<f4').tobytes()))
      else:
        packed = b''.join(_put_varint(int(x) & ((1 << 64) - 1)) for x in arr.reshape(-1))
        feature = _put_field(3, _put_field(1, packed))
    entries += _put_field(1, _put_field(1, key.encode('utf-8')) + _put_field(2, feature))
  return _put_field(1, entries)


# ---- the .spec file ------------------------------------------------------------------------------
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8,
           9: np.int64, 10: np.bool_, 17: np.uint16, 22: np.uint32, 23: np.uint64}


class TensorSpecInfo:
  """shape / dtype (/ bounds) of one leaf of the trajectory spec."""

  def __init__(self, name, shape, dtype, minimum=None, maximum=None):
    self.name, self.shape, self.dtype = name, tuple(shape), np.dtype(dtype)
    self.minimum, self.maximum = minimum, maximum

  def __repr__(self):
    return 'TensorSpecInfo(%r, %s, %s)' % (self.name, self.shape, self.dtype.name)


def _tensor_proto_values(buf: bytes, dtype) -> np.ndarray:
  content, floats = b'', []
  for f, wire, v in _fields(buf):
    if f == 4:                       # tensor_content
      content = v
    elif f == 5:                     # float_val (packed or not)
      floats.extend(np.frombuffer(v, '<f4') if wire == 2
                    else [struct.unpack('<f', struct.pack('<I', v))[0]])
  if content:
    return np.frombuffer(content, dtype=np.dtype(dtype).newbyteorder('<')).astype(dtype)
  return np.array(floats, dtype=dtype)


def _tensor_spec(buf: bytes) -> TensorSpecInfo:
  name, shape, dtype, mn, mx = '', [], np.float32, None, None
  raw_min = raw_max = None
  for f, _, v in _fields(buf):
    if f == 1:
      name = v.decode('utf-8')
    elif f == 2:                     # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
      for g, _, dim in _fields(v):
        if g == 2:
          size = 0
          for h, _, x in _fields(dim):
            if h == 1:
              size = x - (1 << 64) if x >= (1 << 63) else x
          shape.append(size)
    elif f == 3:
      if v not in _DTYPES:
        raise ValueError('unsupported DataType enum %d in spec' % v)
      dtype = _DTYPES[v]
    elif f == 4:
      raw_min = v
    elif f == 5:
      raw_max = v
  if raw_min is not None:
    mn = _tensor_proto_values(raw_min, dtype)
  if raw_max is not None:
    mx = _tensor_proto_values(raw_max, dtype)
  return TensorSpecInfo(name, shape, dtype, mn, mx)


def _structured_value(buf: bytes):
  for f, _, v in _fields(buf):
    if f in (33, 35):                # tensor_spec_value / bounded_tensor_spec_value
      return _tensor_spec(v)
    if f == 54:                      # named_tuple_value { name = 1; repeated PairValue values = 2 }
      out = {}
      for g, _, pair in _fields(v):
        if g == 2:
          key, val = None, None
          for h, _, x in _fields(pair):
            if h == 1:
              key = x.decode('utf-8')
            elif h == 2:
              val = _structured_value(x)
          out[key] = val
      return out
    if f == 53:                      # dict_value { map<string, StructuredValue> fields = 1 }
      out = {}
      for g, _, entry in _fields(v):
        if g == 1:
          key, val = None, None
          for h, _, x in _fields(entry):
            if h == 1:
              key = x.decode('utf-8')
            elif h == 2:
              val = _structured_value(x)
          out[key] = val
      return dict(sorted(out.items()))
    if f in (51, 52):                # list_value / tuple_value { repeated StructuredValue values = 1 }
      return tuple(_structured_value(x) for g, _, x in _fields(v) if g == 1)
    if f == 1:                       # none_value
      return None
  return ()


def parse_spec_file(path: str):
  """`<shard>.spec` -> nest of TensorSpecInfo ({'step_type': ..., 'observation': {...}, ...})."""
  records = list(read_records(path))
  if len(records) != 1:
    raise ValueError('%s: expected one record, found %d' % (path, len(records)))
  return _structured_value(records[0])


def flatten_spec(spec, prefix: str = '') -> Dict[str, TensorSpecInfo]:
  """'/'-joined path -> leaf, in the order tf.nest flattens (dict keys sorted)."""
  out: Dict[str, TensorSpecInfo] = {}
  if isinstance(spec, TensorSpecInfo):
    out[prefix] = spec
  elif isinstance(spec, dict):
    for k, v in spec.items():
      out.update(flatten_spec(v, prefix + '/' + k if prefix else k))
  elif isinstance(spec, tuple):
    for i, v in enumerate(spec):
      out.update(flatten_spec(v, prefix + '/' + str(i) if prefix else str(i)))
  return out


def decode_record(record: bytes, flat_spec: Dict[str, TensorSpecInfo]) -> Dict[str, np.ndarray]:
  """One tf.Example -> {path: array of the spec's shape and dtype}."""
  ex = parse_example(record)
  out = {}
  for path, leaf in flat_spec.items():
    if path not in ex:
      raise KeyError('feature %r of the spec is missing from the record (has %s)'
                     % (path, sorted(ex)))
    raw = ex[path]
    if isinstance(raw, list):        # raw little-endian bytes of the tensor
      if len(raw) != 1:
        raise ValueError('feature %r: expected one bytes value' % path)
      count = int(np.prod(leaf.shape)) if leaf.shape else 1
      if len(raw[0]) != count * leaf.dtype.itemsize:
        raise NotImplementedError(
            'feature %r holds %d bytes for %s %s: compressed image features are not decoded'
            % (path, len(raw[0]), leaf.shape, leaf.dtype.name))
      arr = np.frombuffer(raw[0], dtype=leaf.dtype.newbyteorder('<')).astype(leaf.dtype)
    else:
      arr = np.asarray(raw).astype(leaf.dtype)
    out[path] = arr.reshape(leaf.shape)
  return out


# ---- datasets ---------------------------------------------------------------------------------
STEP_FIRST, STEP_MID, STEP_LAST = 0, 1, 2


def load_shard(path: str, verify: bool = True) -> Dict[str, np.ndarray]:
  """All records of one shard, stacked: {path: [num_records, ...]}."""
  flat = flatten_spec(parse_spec_file(path + '.spec'))
  rows = [decode_record(r, flat) for r in read_records(path, verify=verify)]
  if not rows:
    raise ValueError('%s holds no records' % path)
  return {k: np.stack([r[k] for r in rows]) for k in flat}


def filter_episodes_indices(step_types: np.ndarray) -> np.ndarray:
  """data/dataset.py:29-84: if the window's last FIRST frame is at index s > 0, the frames
  before it belong to another episode and are replaced by copies of frame s."""
  seq_len = step_types.shape[0]
  first = np.nonzero(step_types == STEP_FIRST)[0]
  if first.size == 0 or first[-1] == 0:
    return np.arange(seq_len)
  s = int(first[-1])
  return np.concatenate([np.full(s, s), np.arange(s, seq_len)])


def load_tfrecord_dataset_sequence(paths: Sequence[str], seq_len: int = 1, verify: bool = True):
  """One pass over what `load_tfrecord_dataset_sequence` (data/dataset.py:114-196) streams
  forever: per shard, every window of `seq_len` adjacent records (shift 1; the reference
  windows a repeated shard, so windows wrap from its end to its start), episode-filtered.
  Returns the nest the reference's `get_data` maps it to (get_data.py:57-59):
  observations {key or 'observation': [S, seq_len, ...]} and actions [S, A] of the window's
  last step, plus the window step types [S, seq_len]."""
  obs_out: Dict[str, List[np.ndarray]] = {}
  act_out, st_out = [], []
  specs = None
  for path in paths:
    spec = flatten_spec(parse_spec_file(path + '.spec'))
    sig = {k: (v.shape, v.dtype) for k, v in spec.items()}
    if specs is not None and sig != specs:
      raise ValueError('One or more of the encoding specs do not match.')
    specs = sig
    data = load_shard(path, verify=verify)
    n = data['step_type'].shape[0]
    if n < seq_len:
      continue
    action_keys = [k for k in data if k == 'action' or k.startswith('action/')]
    obs_keys = [k for k in data if k == 'observation' or k.startswith('observation/')]
    for i in range(n):
      idx = (i + np.arange(seq_len)) % n
      idx = idx[filter_episodes_indices(data['step_type'][idx])]
      for k in obs_keys:
        name = k[len('observation/'):] if k.startswith('observation/') else k
        obs_out.setdefault(name, []).append(data[k][idx])
      act = np.concatenate([np.asarray(data[k][idx[-1]], np.float32).reshape(-1)
                            for k in action_keys])
      act_out.append(act)
      st_out.append(data['step_type'][idx])
  observations = {k: np.stack(v) for k, v in obs_out.items()}
  return observations, np.stack(act_out).astype(np.float32), np.stack(st_out)


def episodes_from_shards(paths: Sequence[str], verify: bool = True):
  """Shards -> list of (observation dict of [L, ...], actions [L, A]) split at FIRST frames,
  the layout `ibc_b200.ibc.data.in_memory.windows_from_episodes` consumes."""
  episodes = []
  for path in paths:
    data = load_shard(path, verify=verify)
    starts = list(np.nonzero(data['step_type'] == STEP_FIRST)[0])
    if not starts or starts[0] != 0:
      starts = [0] + starts
    bounds = starts + [data['step_type'].shape[0]]
    action_keys = [k for k in data if k == 'action' or k.startswith('action/')]
    obs_keys = [k for k in data if k == 'observation' or k.startswith('observation/')]
    for a, b in zip(bounds[:-1], bounds[1:]):
      obs = {(k[len('observation/'):] if k.startswith('observation/') else k):
             data[k][a:b].astype(np.float32) for k in obs_keys}
      act = np.concatenate([data[k][a:b].reshape(b - a, -1).astype(np.float32)
                            for k in action_keys], axis=1)
      episodes.append((obs, act))
  return episodes
