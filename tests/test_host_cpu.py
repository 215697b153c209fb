"""CPU-only tests: the C-ABI library loads and exports every symbol the header
declares (no compute calls), host-side logic (schedules, mini-gin, specs), and
the no-CPU-fallback contract."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import mcmc as omcmc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
  from ibc_b200 import build as ibuild
  ibuild.build()
  from ibc_b200 import _lib
  return _lib.load()


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'ibc_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(ibc_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_all_exported(lib):
  names = _declared_symbols()
  assert len(names) >= 25
  for n in names:
    assert hasattr(lib, n), 'libibc_b200.so does not export %s' % n


def test_signature_table_covers_header():
  from ibc_b200 import _lib
  assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_param_count(lib):
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  # C2 267 265; C4 2 136 577; C3-best 8 986 625
  for (o, a, w, d), want in [((20, 2, 128, 16), 267265), ((39, 28, 512, 8), 2136577),
                             ((256, 32, 2048, 2), 8986625)]:
    arch = _lib.MlpArch(o, a, w, d)
    assert lib.ibc_param_count(ctypes.byref(arch)) == want
    assert eng.param_count(o, a, w, d) == want


def test_langevin_stepsizes_host(lib):
  from ibc_b200 import engine as eng
  cfg = eng.langevin_cfg(num_iterations=100)
  got = eng.langevin_stepsizes(cfg)
  want = omcmc.stepsize_table(100)
  assert np.allclose(got, want, rtol=1e-6, atol=0)
  assert np.allclose([got[0], got[1], got[50], got[99]],
                     [0.1, 0.0979902020, 0.0245050505, 1e-5], rtol=2e-7, atol=0)
  cfg = eng.langevin_cfg(num_iterations=5, use_polynomial_rate=False,
                         sampler_stepsize_init=0.1, sampler_stepsize_decay=0.8)
  assert np.allclose(eng.langevin_stepsizes(cfg), [0.1, 0.08, 0.064, 0.0512, 0.04096], rtol=1e-6)


def test_create_without_gpu_fails_loudly(lib):
  if torch.cuda.is_available():
    pytest.skip('has a GPU')
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  with pytest.raises(_lib.IbcError):
    eng.EBMEngine(20, 2, 128, 16)
  arch = _lib.MlpArch(20, 2, 128, 16)
  h = ctypes.c_void_p()
  rc = lib.ibc_create(ctypes.byref(arch), 0, ctypes.byref(h))
  assert rc != 0 and lib.ibc_last_error(None)


def test_odd_depth_rejected(lib):
  from ibc_b200 import _lib
  arch = _lib.MlpArch(20, 2, 128, 3)
  h = ctypes.c_void_p()
  assert lib.ibc_create(ctypes.byref(arch), 0, ctypes.byref(h)) == _lib.IBC_ERR_INVALID
  assert b'even' in lib.ibc_last_error(None)


def test_product_does_not_import_oracle():
  for dirpath, _, files in os.walk(os.path.join(ROOT, 'ibc_b200')):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), (dirpath, f)


def test_gin_lite():
  from ibc_b200 import gin_lite as gin
  gin.clear_config()

  @gin.configurable
  def fn(a, b=2, c='x'):
    return a, b, c

  gin.parse_config("""
    # comment
    fn.b = 1e-3
    fn.c = 'ResNetPreActivation'
    SomethingUnknown.value = True
  """)
  assert fn(1) == (1, 1e-3, 'ResNetPreActivation')
  assert fn(1, b=5) == (1, 5, 'ResNetPreActivation')      # call-site kwargs win
  gin.parse_config_files_and_bindings(None, ['fn.b = 7'])
  assert fn(0)[1] == 7
  gin.clear_config()
  assert fn(0) == (0, 2, 'x')


def test_reference_gin_files_parse():
  """Every EBM gin of the reference uses only `Name.param = literal`."""
  from ibc_b200 import gin_lite as gin
  text = """
train_eval.batch_size = 512
ImplicitBCAgent.num_counter_examples = 256
ImplicitBCAgent.fraction_langevin_samples = 0.0
IbcPolicy.use_dfo = True
IbcPolicy.use_langevin = False
MLPEBM.layers = 'ResNetPreActivation'
MLPEBM.width = 128
MLPEBM.depth = 16
MLPEBM.rate = 0.0
ResNetLayer.normalizer = None
"""
  gin.clear_config()
  gin.parse_config(text)
  assert gin.query_parameter('MLPEBM.depth') == 16
  assert gin.query_parameter('IbcPolicy.use_dfo') is True
  assert gin.query_parameter('ResNetLayer.normalizer') is None
  gin.clear_config()


def test_specs_tile_and_flatten():
  from ibc_b200.ibc import specs
  obs = {'b': torch.arange(4.).reshape(2, 1, 2), 'a': torch.arange(10., 14.).reshape(2, 1, 2)}
  flat = specs.flatten_observation(obs)
  assert flat.tolist() == [[10., 11., 0., 1.], [12., 13., 2., 3.]]     # sorted keys: a then b
  t = specs.tile_batch(torch.tensor([[0.], [1.]]), 3)
  assert t.reshape(-1).tolist() == [0., 0., 0., 1., 1., 1.]


def test_clipped_cd_debug_scalars_match_oracle():
  """The logged scalars of clipped_cd (ibc/losses/ebm_loss.py:138-147) are plain tensor ops in
  the host mirror (ibc_b200/ibc/losses/ebm_loss.py::clipped_cd_debug): against the oracle."""
  import torch
  from ibc_b200.ibc.losses import ebm_loss
  from oracle import losses as olosses
  g = torch.Generator().manual_seed(3)
  pred = torch.randn(6, 10, generator=g, dtype=torch.float64)
  true = (torch.rand(6, 1, 2, generator=g, dtype=torch.float64) - 0.5).expand(6, 9, 2)
  counter = true + (torch.rand(6, 9, 2, generator=g, dtype=torch.float64) - 0.5) * 1.2
  for soft in (False, True):
    _, want = olosses.clipped_cd(pred, counter, true, soft)
    got = ebm_loss.clipped_cd_debug(pred, counter, true, soft)
    assert set(got) == set(want)
    for k in want:
      assert abs(float(got[k]) - float(want[k])) < 1e-12, k
  assert 0.0 < float(want['fraction_outside_threshold']) < 1.0
