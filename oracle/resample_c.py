"""Oracle (test infrastructure): ctypes wrapper of oracle/resample.c."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libibc_oracle.so')
_lib = None


def build():
  subprocess.check_call(['make', '-C', _HERE], stdout=subprocess.DEVNULL)


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(_SO):
      build()
    _lib = ctypes.CDLL(_SO)
    _lib.ibc_oracle_exp64.restype = ctypes.c_double
    _lib.ibc_oracle_exp64.argtypes = [ctypes.c_double]
    _lib.ibc_oracle_resample.restype = ctypes.c_int
    _lib.ibc_oracle_resample.argtypes = [ctypes.c_void_p] * 2 + [
        ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
  return _lib


def exp64(x: float) -> float:
  return lib().ibc_oracle_exp64(float(x))


def resample(logits: np.ndarray, uniforms: np.ndarray):
  """-> (draws [B,count] i32, counts [B,n] i32, sorted_rows [B*count] i64)."""
  logits = np.ascontiguousarray(logits, dtype=np.float32)
  uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
  b, n = logits.shape
  count = uniforms.shape[1]
  draws = np.empty((b, count), np.int32)
  counts = np.empty((b, n), np.int32)
  rows = np.empty(b * count, np.int64)
  rc = lib().ibc_oracle_resample(
      logits.ctypes.data, uniforms.ctypes.data, b, n, count,
      draws.ctypes.data, counts.ctypes.data, rows.ctypes.data)
  assert rc == 0
  return draws, counts, rows
