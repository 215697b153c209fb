"""Builds ibc_b200/lib/libibc_b200.so (the C-ABI of include/ibc_b200.h) with
nvcc for sm_100a.  `python -m ibc_b200.build` or `__graft_entry__.build()`.
The .so is built in-tree so it travels to the GPU box; it is git-ignored."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'lib', 'obj')
LIB = os.path.join(HERE, 'lib', 'libibc_b200.so')
SOURCES = ['api.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'mlp_simt.cu', 'mlp_umma.cu', 'mlp_umma_train.cu', 'mlp_umma_tmem.cu', 'mlp_bwd_wgrad.cu', 'mlp_fwd_halves.cu', 'mlp_wide.cu', 'langevin_tmem.cu', 'langevin_f16.cu', 'peer_allreduce.cu', 'langevin256.cu', 'pixel.cu', 'samplers.cu',
           'losses.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo',
              '-std=c++17', '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  raise RuntimeError('nvcc not found')


def _digest():
  h = hashlib.sha256()
  names = sorted(os.listdir(CSRC)) + ['../../include/ibc_b200.h']
  for n in names:
    with open(os.path.join(CSRC, n), 'rb') as f:
      h.update(n.encode())
      h.update(f.read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def _compile(src, verbose):
  obj = os.path.join(OBJ, src.replace('.cu', '.o'))
  cmd = [_nvcc()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
  if verbose:
    cmd.insert(1, '-Xptxas')
    cmd.insert(2, '-v')
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
  return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OBJ, exist_ok=True)
  stamp = os.path.join(OBJ, 'stamp')
  digest = _digest()
  if (not force and os.path.exists(LIB) and os.path.exists(stamp)
      and open(stamp).read() == digest):
    return LIB
  with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
    results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
  if verbose:
    for _, log in results:
      sys.stderr.write(log)
  objs = [o for o, _ in results]
  cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode',
                                                  'arch=compute_100a,code=sm_100a']
  r = subprocess.run(cmd, capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
  with open(stamp, 'w') as f:
    f.write(digest)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
