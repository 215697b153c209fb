"""ctypes binding of libibc_b200.so (include/ibc_b200.h).

There is NO CPU fallback: if the shared library is missing or the process has
no CUDA device every compute entry point raises.  `oracle/` is never imported
from here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float,
                    c_int, c_int32, c_int64, c_uint32, c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libibc_b200.so')

IBC_OK = 0
IBC_ERR_INVALID = -1
IBC_ERR_UNSUPPORTED = -2
IBC_ERR_CUDA = -3
IBC_ERR_NOT_FINITE = -4

PRECISION_FP32 = 0
PRECISION_TF32 = 1

NORM_NONE, NORM_L1, NORM_L2, NORM_INF = 0, 1, 2, 3
LOSS_BY_NAME = {'info_nce': 0, 'cd': 1, 'clipped_cd': 2, 'soft_clipped_cd': 3}
NORM_BY_NAME = {None: NORM_NONE, '1': NORM_L1, '2': NORM_L2, 'inf': NORM_INF}


class IbcError(RuntimeError):
  pass


class MlpArch(Structure):
  _fields_ = [('obs_dim', c_int32), ('act_dim', c_int32), ('width', c_int32),
              ('depth', c_int32)]


class LangevinCfg(Structure):
  _fields_ = [('num_iterations', c_int32), ('sampler_stepsize_init', c_float),
              ('sampler_stepsize_decay', c_float), ('noise_scale', c_float),
              ('grad_clip', c_float), ('delta_action_clip', c_float),
              ('use_polynomial_rate', c_int32),
              ('sampler_stepsize_final', c_float),
              ('sampler_stepsize_power', c_float), ('grad_norm_type', c_int32),
              ('apply_exp', c_int32)]


class LossCfg(Structure):
  _fields_ = [('softmax_temperature', c_float), ('add_grad_penalty', c_int32),
              ('grad_norm_type', c_int32), ('grad_margin', c_float),
              ('square_grad_penalty', c_int32), ('grad_loss_weight', c_float),
              ('ebm_loss_type', c_int32)]


class PixelArch(Structure):
  _fields_ = [('seq_len', c_int32), ('height', c_int32), ('width', c_int32),
              ('channels', c_int32), ('target_height', c_int32),
              ('target_width', c_int32), ('act_dim', c_int32),
              ('value_width', c_int32), ('value_blocks', c_int32)]


# name -> (restype, argtypes); must list every symbol include/ibc_b200.h declares
SIGNATURES = {
    'ibc_create': (c_int, [POINTER(MlpArch), c_int, POINTER(c_void_p)]),
    'ibc_destroy': (c_int, [c_void_p]),
    'ibc_last_error': (c_char_p, [c_void_p]),
    'ibc_param_count': (c_int64, [POINTER(MlpArch)]),
    'ibc_version': (c_char_p, []),
    'ibc_set_precision': (c_int, [c_void_p, c_int]),
    'ibc_get_precision': (c_int, [c_void_p]),
    'ibc_bind_params': (c_int, [c_void_p, c_void_p, c_void_p]),
    'ibc_params_changed': (c_int, [c_void_p]),
    'ibc_mlp_energy': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                               c_void_p, c_void_p]),
    'ibc_mlp_energy_dact': (c_int, [c_void_p, c_void_p, c_int64, c_void_p,
                                    c_int64, c_void_p, c_void_p, c_void_p]),
    'ibc_resample': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    'ibc_dfo': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                        c_void_p, c_void_p, c_float, c_int32, c_float, c_void_p,
                        c_void_p, c_uint64, c_void_p, c_void_p]),
    'ibc_langevin': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                             c_void_p, c_void_p, POINTER(LangevinCfg), c_void_p,
                             c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ibc_dfo_gather_perturb': (c_int, [c_void_p, c_void_p, c_int64, c_int32,
                                       c_void_p, c_uint64, c_float, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    'ibc_langevin_update': (c_int, [c_void_p, c_void_p, c_int64, c_int32,
                                    c_void_p, c_uint64, c_float, c_float,
                                    c_float, c_float, c_void_p, c_void_p,
                                    c_int32, c_void_p, c_void_p]),
    'ibc_langevin_stepsizes': (c_int, [POINTER(LangevinCfg), POINTER(c_float)]),
    'ibc_softmax_rows': (c_int, [c_void_p, c_int64, c_int64, c_float, c_void_p,
                                 c_void_p]),
    'ibc_uniform_actions': (c_int, [c_void_p, c_int64, c_int32, c_void_p,
                                    c_void_p, c_uint64, c_void_p, c_void_p]),
    'ibc_peer_alloc': (c_int, [c_int64, c_void_p, c_void_p]),
    'ibc_peer_open': (c_int, [c_void_p, c_void_p]),
    'ibc_peer_close': (c_int, [c_void_p, c_int32]),
    'ibc_peer_allreduce_adam': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_void_p,
                                        c_void_p, c_void_p, c_float, c_float, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p]),
    'ibc_greedy_actions': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    'ibc_counter_examples_uniform': (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                             c_uint64, c_void_p, c_void_p, c_void_p]),
    'ibc_info_nce': (c_int, [c_void_p, c_int64, c_int64, c_float, c_float,
                             c_void_p, c_void_p, c_void_p]),
    'ibc_cd_loss': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_float,
                            c_void_p, c_void_p, c_void_p]),
    'ibc_grad_penalty_from_dact': (c_int, [c_void_p, c_int64, c_int64, c_int32,
                                           POINTER(LossCfg), c_float, c_void_p,
                                           c_void_p, c_void_p]),
    'ibc_train_grads': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                POINTER(LossCfg), c_int64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    'ibc_stream_wait_loss': (c_int, [c_void_p, c_void_p, POINTER(c_int32)]),
    'ibc_adam_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_float, c_float, c_float, c_float, c_void_p]),
    'ibc_pixel_create': (c_int, [POINTER(PixelArch), c_int, POINTER(c_void_p)]),
    'ibc_pixel_destroy': (c_int, [c_void_p]),
    'ibc_pixel_last_error': (c_char_p, [c_void_p]),
    'ibc_pixel_param_count': (c_int64, [POINTER(PixelArch)]),
    'ibc_pixel_bind_params': (c_int, [c_void_p, c_void_p]),
    'ibc_pixel_set_precision': (c_int, [c_void_p, c_int]),
    'ibc_pixel_get_precision': (c_int, [c_void_p]),
    'ibc_pixel_kernel_status': (c_int, [c_void_p, POINTER(c_int32)]),
    'ibc_pixel_encode': (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p]),
    'ibc_pixel_energy': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                 c_void_p]),
    'ibc_pixel_dfo': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                              c_void_p, c_float, c_int32, c_float, c_void_p, c_void_p,
                              c_uint64, c_void_p, c_void_p]),
    'ibc_pixel_energy_dact': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                      c_void_p, c_void_p]),
    'ibc_pixel_langevin': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                   c_void_p, POINTER(LangevinCfg), c_void_p, c_uint64,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    'ibc_pixel_train_grads': (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int64,
                                      POINTER(LossCfg), c_int64, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    'ibc_pixel_train_begin': (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p,
                                      c_void_p]),
    'ibc_pixel_train_finish': (c_int, [c_void_p, c_void_p, POINTER(LossCfg), c_int64, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    'ibc_kernel_status': (c_int, [c_void_p, POINTER(c_int32)]),
    'ibc_set_profiling': (c_int, [c_void_p, c_int32]),
    'ibc_get_train_timings': (c_int, [c_void_p, POINTER(c_float)]),
    'ibc_launch_count': (c_int64, []),
    'ibc_reset_launch_count': (None, []),
    'ibc_umma_selftest': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                  c_void_p, POINTER(c_int32), c_void_p]),
    'ibc_umma_f16_selftest': (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_uint32, c_uint32,
                                      c_uint32, c_uint32, c_uint32, c_void_p, POINTER(c_int32),
                                      c_void_p]),
    'ibc_gemm_tf32': (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p,
                              c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                              c_int32, c_int32, c_int32, POINTER(c_int32), c_void_p]),
    'ibc_conv3x3_tf32': (c_int, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_int32,
                                 c_void_p, c_int32, c_int32, c_void_p, c_void_p, POINTER(c_int32),
                                 c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
  """Loads the shared library (never builds it silently on a GPU box: run
  `python -m ibc_b200.build` or `__graft_entry__.build()` first)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise IbcError(
        'libibc_b200.so is not built (%s). Run `python -m ibc_b200.build`; '
        'there is no CPU fallback.' % LIB_PATH)
  lib = ctypes.CDLL(LIB_PATH)
  for name, (restype, argtypes) in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
  _lib = lib
  return lib


_EXC = {IBC_ERR_INVALID: ValueError, IBC_ERR_UNSUPPORTED: NotImplementedError,
        IBC_ERR_CUDA: IbcError, IBC_ERR_NOT_FINITE: FloatingPointError}


def check(rc: int, handle=None, pixel=False):
  if rc == IBC_OK:
    return
  msg = (load().ibc_pixel_last_error if pixel else load().ibc_last_error)(handle)
  msg = msg.decode('utf-8', 'replace') if msg else 'error %d' % rc
  raise _EXC.get(rc, IbcError)(msg)


def launch_count() -> int:
  return int(load().ibc_launch_count())


def reset_launch_count():
  load().ibc_reset_launch_count()
