"""Benchmark of the IBC EBM hot path on B200 (contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): block-pushing-states MLP EBM
(pushing_states/mlp_ebm.gin: obs 2x10, act 2, width 128, depth 16, batch 512,
256 uniform counter-examples, InfoNCE, no gradient penalty).  A step is one
`ImplicitBCAgent.train` on one synthetic batch: negatives -> B*(N+1)-row
energy forward -> InfoNCE -> backward -> (NCCL all-reduce) -> Adam.
metric = energy evaluations per second over all ranks.

Extra blocks in the JSON line: `roofline` (dominant kernel of the step),
`roofline_energy_fwd` (the tcgen05 fused forward kernel timed alone),
`dfo_inference` (IbcPolicy actions/s at N=16384, DFO 3 iterations),
`cpu_baseline` (the oracle = CPU restatement of the reference, timed on this
box's host cores; the TF reference itself cannot be installed offline).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import numpy as np
import torch

CFG = dict(name='pushing_states/mlp_ebm.gin (C2)', obs_dim=20, seq=2, act_dim=2, width=128,
           depth=16, batch=512, num_counter_examples=256, lr=1e-3, decay_steps=100,
           dfo_samples=16384, dfo_iterations=3)


def workload_config(c=CFG):
  """`config` of the JSON line -- the SAME dict in the b200 and the reference arm."""
  rows = c['batch'] * (c['num_counter_examples'] + 1)
  return {'workload': c['name'], 'batch_per_gpu': c['batch'],
          'num_counter_examples': c['num_counter_examples'], 'rows_per_step_per_gpu': rows,
          'obs_dim': c['obs_dim'], 'act_dim': c['act_dim'], 'width': c['width'], 'depth': c['depth'],
          'l2': 'per-step saved activations (%.2f GB fp16 + fp32 residual) exceed the 126 MB L2; '
                'the timed steps cycle over 4 resident batches' % (
                    (c['depth'] * 2 + 4) * rows * c['width'] / 1e9)}


def measure_tf32_peak(device, seconds=0.4):
  """the tf32 tensor peak measured on the box -- torch.matmul (cuBLAS)
  8192^3 with tf32 operands, best of 10 (burst) and a back-to-back loop (sustained)."""
  try:
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device=device)
    b = torch.randn(n, n, device=device)
    for _ in range(3):
      a @ b
    best = 1e9
    for _ in range(10):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
      best = min(best, e0.elapsed_time(e1) * 1e-3)
    reps = max(3, int(seconds / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      a @ b
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) * 1e-3 / reps
    torch.backends.cuda.matmul.allow_tf32 = prev
    del a, b
    return {'burst': 2.0 * n ** 3 / best / 1e12, 'sustained': 2.0 * n ** 3 / sustained / 1e12,
            'how': 'torch.matmul fp32 inputs, allow_tf32, 8192^3: best of 10 / %d back to back' % reps}
  except Exception as ex:   # noqa: BLE001
    return {'burst': None, 'sustained': None, 'how': 'failed: %s' % str(ex)[:100]}


def fwd_flops_per_row(c=CFG):
  d_in = c['obs_dim'] + c['act_dim']
  return 2 * (d_in * c['width'] + c['depth'] * c['width'] ** 2 + c['width'])


def measured_peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    try:
      d = json.load(open(p))
      return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sustained=d.get(
          'bf16_tflops_sustained', d['bf16_tflops']), source='measured')
    except Exception:
      pass
  return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
  Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.rows, self.proc, self.first = index, [], None, 0

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
           '--format=csv,noheader,nounits', '-lms', '100'],
          stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def mark(self):
    """Rows printed so far (start-up, warm-up) are not part of the report."""
    self.first = len(self.rows)

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([x.strip() for x in line.split(',')])

  def stop(self):
    if not self.proc:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, mx, reasons = [], [], set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    for r in self.rows[self.first:]:
      try:
        sm.append(float(r[0])); mx.append(float(r[1]))
        for nm, v in zip(names, r[2:6]):
          if v.lower().startswith('active'):
            reasons.add(nm)
      except Exception:
        continue
    return {'sm_mhz': float(np.median(sm)) if sm else None,
            'sm_max_mhz': max(mx) if mx else None, 'reasons': sorted(reasons),
            'samples': len(sm)}


def synthetic_batch(c, batch, seed, device=None, pinned=False):
  """obs ~ N(0,1) (z-scored), true actions ~ U[-1,1] (min-max normalised)."""
  g = torch.Generator().manual_seed(seed)
  half = c['obs_dim'] // c['seq'] // 2
  obs = {'block_translation': torch.randn(batch, c['seq'], half, generator=g),
         'effector_translation': torch.randn(batch, c['seq'], c['obs_dim'] // c['seq'] - half,
                                             generator=g)}
  act = torch.rand(batch, c['act_dim'], generator=g) * 2 - 1
  if pinned:
    obs = {k: v.pin_memory() for k, v in obs.items()}
    act = act.pin_memory()
  if device is not None:
    obs = {k: v.to(device) for k, v in obs.items()}
    act = act.to(device)
  return obs, act


def build_agent(c, device, seed=1):
  from ibc_b200.ibc import specs
  from ibc_b200.ibc.agents import ibc_agent
  from ibc_b200.ibc.train import get_agent
  from ibc_b200.networks.mlp_ebm import MLPEBM
  half = c['obs_dim'] // c['seq'] // 2
  obs_spec = {'block_translation': specs.TensorSpec((c['seq'], half)),
              'effector_translation': specs.TensorSpec((c['seq'], c['obs_dim'] // c['seq'] - half))}
  act_spec = specs.BoundedTensorSpec((c['act_dim'],), minimum=-1.0, maximum=1.0)
  sampling_spec = specs.BoundedTensorSpec((c['act_dim'],), minimum=-1.1, maximum=1.1)
  net = MLPEBM((obs_spec, act_spec), specs.TensorSpec((1,)), width=c['width'], depth=c['depth'],
               rate=0.0, layers='ResNetPreActivation', device=device, seed=seed)
  opt = get_agent.Adam(get_agent.ExponentialDecay(c['lr'], c['decay_steps'], 0.99))
  agent = ibc_agent.ImplicitBCAgent(
      time_step_spec=obs_spec, action_spec=act_spec, action_sampling_spec=sampling_spec,
      cloning_network=net, optimizer=opt, num_counter_examples=c['num_counter_examples'],
      fraction_langevin_samples=0.0, add_grad_penalty=False, compute_mse=True,
      return_full_chain=True)
  policy = None
  from ibc_b200.ibc.agents import ibc_policy
  policy = ibc_policy.GreedyPolicy(ibc_policy.IbcPolicy(
      obs_spec, act_spec, sampling_spec, net, num_action_samples=c['dfo_samples'], use_dfo=True,
      use_langevin=False))
  return net, agent, policy


def timed(fn, steps, barrier):
  """CUDA-event time of `steps` calls on the current stream, barrier + sync on both sides."""
  barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(steps):
    fn(i)
  e1.record()
  torch.cuda.synchronize()
  barrier()
  return e0.elapsed_time(e1) / 1e3


def run_b200(args):
  import torch.distributed as dist
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  from ibc_b200.ibc import specs
  c = CFG
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  device = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)

  def barrier():
    if world > 1:
      dist.barrier()

  def max_over_ranks(x):
    if world == 1:
      return x
    t = torch.tensor([x], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  net, agent, policy = build_agent(c, device)
  # weak scaling: every rank trains on its own batch of 512 (per-replica batch fixed)
  b, n = c['batch'], c['num_counter_examples']
  rows = b * (n + 1)
  nb = 4
  dev_batches = [synthetic_batch(c, b, 100 + rank * 17 + i, device=device) for i in range(nb)]
  host_batches = [synthetic_batch(c, b, 100 + rank * 17 + i, pinned=True) for i in range(nb)]
  total = args.steps + args.warmup

  def step_dev(i):
    agent.train((dev_batches[i % nb], ()))

  # W warm-up steps as asked, and never fewer than 300: the step's temporaries pass between two
  # streams (record_stream), so the caching allocator keeps growing -- cudaMalloc, a device-wide
  # stall, on whichever rank needs it, with every other rank waiting in the gradient sum -- for
  # a few hundred steps (4 GPUs, 200 timed steps: 0.533 ms/step after 20 warm-up steps, 0.479
  # after 400; `sustained` 0.477-0.484 either way)
  warm_run = max(args.warmup, 300)
  # nvidia-smi starts BEFORE the warm-up (its start-up -- NVML initialisation inside the driver --
  # stalled the first launches of a 20-step timed region by 0.8 ms); only the rows it prints
  # from the start of the timed region on are used
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  # the interpreter's start-up heap (torch, numpy: ~10^6 objects) out of the cyclic collector's
  # way: a full collection over it is a pause of tens of milliseconds on one rank, which every
  # other rank of a data-parallel run then waits for.  The collector itself stays on.  (Before
  # the warm-up, not after it: a host pause in front of the timed region idles the GPU.)
  import gc
  gc.collect()
  gc.freeze()
  for i in range(warm_run):
    step_dev(i)
  torch.cuda.synchronize()
  sampler.mark()
  _lib.reset_launch_count()
  t_dev = max_over_ranks(timed(step_dev, args.steps, barrier))
  launches = _lib.launch_count()
  # the same step for >= 0.4 s so that nvidia-smi (100 ms period) samples the clocks under
  # load even when K is small (the driver passes --steps 20 = 15 ms)
  sus_steps = max(args.steps, int(0.4 / max(t_dev / args.steps, 1e-5)))
  t_sus = max_over_ranks(timed(step_dev, sus_steps, barrier))
  clocks = sampler.stop() if rank == 0 else None

  # end to end: pinned host buffers -> H2D -> step -> loss back on the host
  losses = []

  def step_e2e(i):
    info = agent.train((host_batches[i % nb], ()))
    losses.append(float(info.loss))          # D2H read of the step's result

  for i in range(min(3, args.warmup)):
    step_e2e(i)
  barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for i in range(args.steps):
    step_e2e(i)
  torch.cuda.synchronize()
  t_e2e = max_over_ranks(time.perf_counter() - t0)
  h2d = sum(v.numel() * 4 for v in host_batches[0][0].values()) + host_batches[0][1].numel() * 4

  # per-kernel CUDA-event timings of the tensor-core training step (same steps,
  # profiling on: events recorded on the launching stream around each phase)
  kt = None
  try:
    net.engine.set_profiling(True)
    acc = np.zeros(5)
    nprof = min(args.steps, 50)
    for i in range(nprof):
      step_dev(i)
      acc += np.array(net.engine.train_timings())
    kt = (acc / nprof).tolist()
  except Exception as ex:   # noqa: BLE001  (fp32 path has no phase events)
    kt = None
  finally:
    net.engine.set_profiling(False)

  # the same step on the exact fp32 FMA path (IBC_PRECISION_FP32)
  t_fp32 = None
  try:
    net.engine.set_precision(_lib.PRECISION_FP32)
    for i in range(2):
      step_dev(i)
    t_fp32 = max_over_ranks(timed(step_dev, 5, barrier)) / 5
  except Exception:   # noqa: BLE001
    t_fp32 = None
  finally:
    net.engine.set_precision(_lib.PRECISION_TF32)

  # the fused tcgen05 energy forward timed alone on the step's rows
  obs_flat = net._flat_obs(dev_batches[0][0])
  acts = torch.rand(rows, c['act_dim'], device=device) * 2.2 - 1.1
  peaks = measured_peaks()
  # (measured LAST -- a 0.4 s cuBLAS loop pulls the clocks down for whatever runs right after it
  # -- and substituted into the roofline blocks below; see the end of this function)
  tf32_meas = {'burst': None, 'sustained': None, 'how': ''}
  # tensor roofline denominator: the tf32 cuBLAS figure measured here,
  # else half the driver-measured bf16 burst (the nominal tf32 : bf16 ratio)
  tf32_peak = tf32_meas['burst'] if tf32_meas.get('burst') else peaks['bf16'] / 2.0
  tf32_peak_source = ('tf32 cuBLAS 8192^3 measured in this run' if tf32_meas.get('burst')
                      else '%s bf16 burst / 2 (tf32 nominal ratio)' % peaks['source'])
  fwd = {}
  for prec, label in ((_lib.PRECISION_TF32, 'tf32'), (_lib.PRECISION_FP32, 'fp32')):
    try:
      net.engine.set_precision(prec)
      for _ in range(3):
        net.engine.energy(obs_flat, acts, n + 1)
      reps = 20 if prec == _lib.PRECISION_TF32 else 5
      t = timed(lambda i: net.engine.energy(obs_flat, acts, n + 1), reps, lambda: None) / reps
      fwd[label] = t
    except Exception as ex:   # noqa: BLE001
      fwd[label] = None
      fwd[label + '_error'] = str(ex)[:200]
  net.engine.set_precision(_lib.PRECISION_TF32)
  status = net.engine.kernel_status()

  # DFO policy inference (B=1, N=16384, 3 iterations), actions/s
  ts_obs = {k: v[:1] for k, v in dev_batches[0][0].items()}
  time_step = specs.restart(ts_obs)
  for _ in range(3):
    policy.action(time_step)
  reps = 20
  t_act = timed(lambda i: policy.action(time_step), reps, lambda: None) / reps

  # batched policy inference sharded by observation (north_star): every rank
  # evaluates 32 observations x 16384 candidates (DFO, 3 iterations) of a 32 * N batch and the
  # actions are gathered (A floats per observation); weak scaling, no data-path collective
  from ibc_b200 import parallel
  obs_per_gpu = 32
  big, _ = synthetic_batch(c, obs_per_gpu * world, 777, device=device)
  big_ts = specs.restart(big)
  for _ in range(2):
    parallel.sharded_policy_action(policy, big_ts)
  reps = 5
  t_shard = max_over_ranks(timed(lambda i: parallel.sharded_policy_action(policy, big_ts), reps,
                                 barrier)) / reps

  # secondary configs: Langevin chains of C3 / pushing_states-langevin / C4, C1 training
  # gradients and the PixelEBM paths on rank 0; the C3 / C4 train steps through
  # ImplicitBCAgent.train on EVERY rank (they all-reduce their gradients)
  secondary = {}
  try:
    secondary = secondary_configs(device, rank, world)
  except Exception as ex:   # noqa: BLE001
    secondary = {'error': str(ex)[:300]}

  result = None
  if rank == 0:
    evals = rows * world * args.steps
    step_flops = 3.0 * rows * fwd_flops_per_row()
    ms = t_dev / args.steps * 1e3
    achieved = step_flops / (t_dev / args.steps) / 1e12
    result = {
        'metric': 'ebm_train_energy_evals_per_s', 'value': evals / t_dev, 'unit': 'energy evals/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'warmup_steps_run': warm_run,
        'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'tf32 (tcgen05 kind::tf32 operands, fp32 accumulate; first/last layer and loss in fp32)', 'data': 'synthetic',
        'config': workload_config(c),
        'train_steps_per_s': args.steps / t_dev,
        'sustained': {'steps': sus_steps, 'ms_per_step': t_sus / sus_steps * 1e3,
                      'value': rows * world * sus_steps / t_sus,
                      'note': 'same step, long enough for the clock samples below'},
        'fp32_exact_step': ({'ms_per_step': t_fp32 * 1e3, 'value': rows * world / t_fp32,
                             'note': 'IBC_PRECISION_FP32: every contraction on fp32 FMA kernels'}
                            if t_fp32 else None),
        'tf32_tflops_measured': tf32_meas,
        'inference_sharded': {'actions_per_s': obs_per_gpu * world / t_shard,
                              'ms_per_batch': t_shard * 1e3, 'observations': obs_per_gpu * world,
                              'observations_per_gpu': obs_per_gpu,
                              'num_action_samples': c['dfo_samples'], 'iterations': c['dfo_iterations'],
                              'collective': 'none on the data path; one all_gather of A floats per observation',
                              'scaling': 'weak'},
        'e2e': {'value': evals / t_e2e, 'unit': 'energy evals/s', 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': 4, 'ms_per_step': t_e2e / args.steps * 1e3,
                'final_loss': losses[-1] if losses else None},
        'gpu_launches': int(launches),
        'gradient_reduction': (None if world == 1 else
                               'ibc_peer_allreduce_adam: sum over NVLink peer memory + Adam in one launch'
                               if getattr(agent, '_peer', None) is not None else
                               'torch.distributed.all_reduce (NCCL) + ibc_adam_step'),
        'clocks': clocks,
        'roofline': None,
        'roofline_step': {'kernel': 'whole train step (all kernels)', 'bound': 'tensor',
                          'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s',
                          'frac': achieved / tf32_peak, 'flops_per_step': step_flops,
                          'peak_source': tf32_peak_source},
        'roofline_energy_fwd': None,
        'dfo_inference': {'actions_per_s': 1.0 / t_act, 'ms_per_action': t_act * 1e3,
                          'num_action_samples': c['dfo_samples'],
                          'iterations': c['dfo_iterations']},
        'kernel_status': status,
        'secondary': secondary,
    }
    if fwd.get('tf32'):
      fl = rows * fwd_flops_per_row()
      a = fl / fwd['tf32'] / 1e12
      result['roofline_energy_fwd'] = {
          'kernel': 'mlp_fwd_halves_kernel<0> (tcgen05 kind::tf32, activations in tensor memory)', 'bound': 'tensor',
          'achieved': a, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': a / tf32_peak,
          'peak_source': tf32_peak_source,
          'traffic': _ncu_traffic().get('forward_nosave'), 'ms_per_launch': fwd['tf32'] * 1e3, 'rows': rows,
          'fp32_path_ms': fwd['fp32'] * 1e3 if fwd.get('fp32') else None}
    else:
      result['roofline_energy_fwd'] = {k: v for k, v in fwd.items()}
    if kt is not None:
      # dominant kernel of the step = the one with the largest measured share.  Algorithmic
      # work per launch: the fused forward = rows x fwd FLOPs; the
      # block-major backward = backward-data + backward-weights = 2 x rows x fwd FLOPs
      # (SURVEY.md 8d canonical count).  Algorithmic HBM bytes: the forward writes, and the
      # backward reads, D fp16 operand tiles + the fp32 final residual + D mask bit-planes per
      # row; the backward also writes g0 (fp32) -- its parked gradients are L2 traffic.
      fflops = rows * fwd_flops_per_row()
      names = ['mlp_fwd_halves_kernel<2, true> (fused forward in kind::f16, fp16 operand tiles + ReLU masks saved)',
               'info_nce_kernel',
               'mlp_bwd_wgrad_tmem_kernel (block-major reverse chain + weight gradients)',
               '(weight gradients: inside the backward kernel)',
               'first-layer fp32 kernels']
      flops = [fflops, 0.0, 2.0 * fflops, 0.0, 0.0]
      tile16 = rows * c['width'] * 2.0
      tile32 = rows * c['width'] * 4.0
      mask_bytes = rows * c['width'] / 8.0
      saved = c['depth'] * (tile16 + mask_bytes) + tile32
      hbm = [saved, 0.0, saved + tile32, 0.0, tile32]
      top = int(np.argmax(kt))
      keys = ['forward', 'info_nce', 'backward', 'wgrad', 'first_layer']
      result['kernel_ms'] = dict(zip(keys, kt))
      traffic = _ncu_traffic().get(keys[top])
      tfl = flops[top] / (kt[top] * 1e-3) / 1e12 if flops[top] else 0.0
      gbs = hbm[top] / (kt[top] * 1e-3) / 1e9
      result['roofline'] = {
          'kernel': names[top], 'bound': 'tensor', 'achieved': tfl, 'peak': tf32_peak,
          'unit': 'TFLOP/s', 'frac': tfl / tf32_peak, 'traffic': traffic,
          'ms_per_launch': kt[top], 'share_of_step': kt[top] / sum(kt),
          'algorithmic_flops_per_launch': flops[top],
          'hbm_algorithmic_GBs': gbs, 'hbm_frac_of_measured': gbs / peaks['hbm'],
          'algorithmic_bytes_per_launch': hbm[top],
          'peak_source': tf32_peak_source,
          'note': 'chain MMAs kind::tf32, weight-gradient MMAs kind::f16 (same 10-bit mantissa); '
                  'both counted against the tf32 peak'}
      result['roofline_kernels'] = [
          {'kernel': names[i], 'ms': kt[i], 'tflops': flops[i] / (kt[i] * 1e-3) / 1e12 if flops[i] else None,
           'frac_of_tf32_peak': (flops[i] / (kt[i] * 1e-3) / 1e12 / tf32_peak) if flops[i] else None,
           'hbm_GBs': hbm[i] / (kt[i] * 1e-3) / 1e9 if hbm[i] else None} for i in range(5)]
    else:
      result['roofline'] = dict(result['roofline_step'])
      result['roofline']['kernel'] = 'train step: per-layer fp32 FMA gemm_kernel launches'
      result['roofline']['traffic'] = None
    if world == 1:
      # rank 0 at N = 1 only: with more ranks the others spin in the closing barrier on the
      # same host cores and the figure means nothing (63 k instead of 300 k evals/s at N = 8)
      result['cpu_baseline'] = cpu_baseline(steps=2, warmup=1)
      try:
        result['cpu_baselines_secondary'] = cpu_baselines_secondary()
      except Exception as ex:   # noqa: BLE001
        result['cpu_baselines_secondary'] = {'error': str(ex)[:200]}
  if rank == 0 and result is not None:
    tf32_meas = measure_tf32_peak(device)
    result['tf32_tflops_measured'] = tf32_meas
    if tf32_meas.get('burst'):
      pk = tf32_meas['burst']
      src = 'tf32 cuBLAS 8192^3 measured at the end of this run'
      for key in ('roofline', 'roofline_step', 'roofline_energy_fwd'):
        blk = result.get(key)
        if isinstance(blk, dict) and blk.get('unit') == 'TFLOP/s' and blk.get('achieved') is not None:
          blk['peak'], blk['frac'], blk['peak_source'] = pk, blk['achieved'] / pk, src
      for blk in result.get('roofline_kernels') or []:
        if blk.get('tflops'):
          blk['frac_of_tf32_peak'] = blk['tflops'] / pk
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return result


def _ncu_traffic():
  """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step's kernels,
  from this round's committed `ncu --set full` capture of the same step
  (profiles/r2_dram_traffic.json, written by scripts/ncu_summary.py)."""
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'r2_dram_traffic.json')
  try:
    with open(path) as f:
      return json.load(f)
  except (OSError, ValueError):
    return {}


def secondary_configs(device, rank=0, world=1):
  """Langevin sampler throughput (C3: particle 32-D, W=256 D=2, obs 256, act 32, B*N=4096,
  K=100; pushing_states/mlp_ebm_langevin.gin: W=128 D=16, B*N=4096, K=100) and a PixelEBM
  train step (C5 shapes, B=32 of 128 frames pairs)."""
  from ibc_b200 import engine as eng
  out = {}
  # Full train steps of the Langevin configs through ImplicitBCAgent.train (the reference's
  # call): uniform seeds -> 100-step Langevin chain on B*N = 4096 rows -> InfoNCE + gradient
  # penalty forward / second-order backward on B*(N+1) rows -> Adam (energy
  # evals per step = K*B*N + 2*B*(N+1))
  from ibc_b200 import gin_lite as gin
  from ibc_b200.ibc import specs
  from ibc_b200.ibc.agents import ibc_agent
  from ibc_b200.ibc.train import get_agent
  from ibc_b200.networks.mlp_ebm import MLPEBM
  import torch.distributed as dist
  # These run on EVERY rank: ImplicitBCAgent.train all-reduces its gradients (weak scaling,
  # per-replica batch 512); the time is the max over ranks.  Everything after them runs on
  # rank 0 only, and the single-process agent steps (PixelEBM) only without a process group.
  solo = world == 1

  def max_ranks(x):
    if world == 1:
      return x
    t_ = torch.tensor([x], device=device, dtype=torch.float64)
    dist.all_reduce(t_, op=dist.ReduceOp.MAX)
    return float(t_.item())

  for name, (o, a, w, d), extra in (
      ('C3_particle32_train_step', (256, 32, 256, 2), ''),
      ('C4_d4rl_door_train_step', (39, 28, 512, 8),
       'langevin_actions_given_obs.sampler_stepsize_init = 0.5\n'
       'langevin_actions_given_obs.delta_action_clip = 0.5\n'
       'langevin_actions_given_obs.noise_scale = 0.5\n')):
    gin.clear_config()
    gin.parse_config('langevin_actions_given_obs.num_iterations = 100\n' + extra)
    try:
      b, n, k = 512, 8, 100
      obs_spec = {'obs': specs.TensorSpec((1, o))}
      act_spec = specs.BoundedTensorSpec((a,), minimum=-1.0, maximum=1.0)
      samp = specs.BoundedTensorSpec((a,), minimum=-1.1, maximum=1.1)
      net = MLPEBM((obs_spec, act_spec), specs.TensorSpec((1,)), width=w, depth=d, rate=0.0,
                   layers='ResNetPreActivation', seed=1)
      opt = get_agent.Adam(get_agent.ExponentialDecay(1e-3, 100, 0.99))
      agent = ibc_agent.ImplicitBCAgent(obs_spec, act_spec, samp, net, opt, num_counter_examples=n,
                                        fraction_langevin_samples=1.0, add_grad_penalty=True,
                                        run_full_chain_under_gradient=True)
      g = torch.Generator().manual_seed(rank)
      exp = (({'obs': torch.randn(b, 1, o, generator=g).to(device)},
              (torch.rand(b, a, generator=g) * 2 - 1).to(device)), ())
      for _ in range(3):
        agent.train(exp)
      t = max_ranks(timed(lambda i: agent.train(exp), 5, lambda: None) / 5)
      evals = (k * b * n + 2 * b * (n + 1)) * world
      fwd = 2.0 * ((o + a) * w + d * w * w + w)
      flops = (k * b * n * 2 + b * (n + 1) * 7) * fwd * world
      out[name] = {'ms_per_step': t * 1e3, 'batch': b, 'num_counter_examples': n,
                   'langevin_iterations': k, 'energy_evals_per_s': evals / t,
                   'tflops': flops / t * 1e-12, 'n_gpus': world, 'precision': 'tf32',
                   'kernel_status': net.engine.kernel_status()}
    finally:
      gin.clear_config()
  if rank != 0:
    return out
  for name, (o, a, w, d) in (('C3_particle32_langevin', (256, 32, 256, 2)),
                             ('pushing_states_langevin', (20, 2, 128, 16)),
                             ('C4_d4rl_door_langevin', (39, 28, 512, 8))):
    e = eng.EBMEngine(o, a, w, d, device)
    params = (torch.randn(e.param_count, device=device) * 0.05).contiguous()
    e.bind_params(params)
    b, n, k = 512, 8, 100
    obs = torch.randn(b, o, device=device)
    mn, mx = torch.full((a,), -1.1, device=device), torch.full((a,), 1.1, device=device)
    cfg = eng.langevin_cfg(num_iterations=k)
    acts = torch.rand(b * n, a, device=device) * 2.2 - 1.1
    for w in range(3):      # the second call records the chain's CUDA graph, later calls replay it
      e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=1 + w)
    t = timed(lambda i: e.langevin(obs, acts.clone(), n, mn, mx, cfg, seed=10 + i), 5, lambda: None) / 5
    out[name] = {'ms_per_chain': t * 1e3, 'rows': b * n, 'iterations': k,
                 'energy_grad_evals_per_s': b * n * k / t, 'precision': 'tf32',
                 'kernel_status': e.kernel_status()}
  # C1 (particle 2-D, DFO-style training: W=256 D=2, B=512, N=256 uniform negatives, no
  # gradient penalty): loss forward + backward + weight gradients on 131 584 rows; width 256
  # trains through the per-layer passes on the general tcgen05 GEMM
  from ibc_b200 import _lib as _l
  e = eng.EBMEngine(16, 2, 256, 2, device)
  e.bind_params((torch.randn(e.param_count, device=device) * 0.05).contiguous())
  b, n = 512, 256
  obs = torch.randn(b, 16, device=device)
  acts = torch.rand(b * (n + 1), 2, device=device) * 2.2 - 1.1
  cfg = eng.loss_cfg(1.0, False)
  fwd = 2.0 * (18 * 256 + 2 * 256 * 256 + 256)
  for prec, key in ((_l.PRECISION_TF32, 'C1_particle2d_train_grads'),
                    (_l.PRECISION_FP32, 'C1_particle2d_train_grads_fp32')):
    e.set_precision(prec)
    for _ in range(2):
      e.train_grads(obs, acts, n, cfg, b)
    t = timed(lambda i: e.train_grads(obs, acts, n, cfg, b), 5, lambda: None) / 5
    out[key] = {'ms_per_step': t * 1e3, 'rows': b * (n + 1), 'energy_evals_per_s': b * (n + 1) / t,
                'tflops': 3.0 * b * (n + 1) * fwd / t * 1e-12,
                'precision': 'tf32' if prec == _l.PRECISION_TF32 else 'fp32',
                'kernel_status': e.kernel_status()}
  # C5 PixelEBM (pushing_pixels/pixel_ebm_best.gin shapes): late-fusion train step at the
  # gin's batch 128 with 256 counter-examples, tf32 tensor-core path and the fp32 path
  from ibc_b200 import _lib
  pe = eng.PixelEngine(2, 240, 320, 3, 180, 240, 2, 1024, 1, device)
  pp = (torch.randn(pe.param_count, device=device) * 0.05).contiguous()
  pe.bind_params(pp)
  b, n = 128, 256
  img = torch.randint(0, 256, (b, 2, 240, 320, 3), device=device, dtype=torch.uint8)
  acts = torch.rand(b * (n + 1), 2, device=device) * 2 - 1
  cfg = eng.loss_cfg(1.0, False)
  # canonical FLOPs: encoder 1.335 G / image, value head 1 710 080 / row, x 3
  flops = 3.0 * (b * 1.335e9 + b * (n + 1) * 1710080.0)
  for prec, key, label in ((_lib.PRECISION_TF32, 'C5_pixel_ebm_train_grads',
                            'tf32 (tcgen05 implicit-GEMM convolutions + value head, fp32 accumulate)'),
                           (_lib.PRECISION_FP32, 'C5_pixel_ebm_train_grads_fp32',
                            'fp32 (im2col + FMA GEMM)')):
    pe.set_precision(prec)
    pe.train_grads(img, acts, n, cfg, b)
    t = timed(lambda i: pe.train_grads(img, acts, n, cfg, b), 3, lambda: None) / 3
    out[key] = {'ms_per_step': t * 1e3, 'images': b, 'rows': b * (n + 1), 'images_per_s': b / t,
                'energy_evals_per_s': b * (n + 1) / t, 'tflops': flops / t * 1e-12,
                'precision': label, 'kernel_status': pe.kernel_status()}
  # C5 whole train step through the reference's call (late-fusion ImplicitBCAgent.train on a
  # PixelEBM: uniform negatives, encode once, tile, InfoNCE, backward, Adam), frames resident
  # on the device; single-process runs only (see above)
  if solo:
    from ibc_b200.networks.pixel_ebm import PixelEBM
    gin.clear_config()
    gin.parse_config('DenseResnetValue.width = 1024\nDenseResnetValue.num_blocks = 1')
    try:
      p_obs = {'rgb': specs.TensorSpec((2, 240, 320, 3), dtype=torch.uint8)}
      p_act = specs.BoundedTensorSpec((2,), minimum=-1.0, maximum=1.0)
      p_samp = specs.BoundedTensorSpec((2,), minimum=-1.1, maximum=1.1)
      pnet = PixelEBM((p_obs, p_act), specs.TensorSpec((1,)), 'ConvMaxpoolEncoder',
                      'DenseResnetValue', target_height=180, target_width=240, seed=0)
      pagent = ibc_agent.ImplicitBCAgent(p_obs, p_act, p_samp, pnet, get_agent.Adam(1e-3),
                                         num_counter_examples=n, fraction_langevin_samples=0.0,
                                         late_fusion=True, add_grad_penalty=False)
      pexp = (({'rgb': img}, torch.rand(b, 2, device=device) * 2 - 1), ())
      for _ in range(2):
        pagent.train(pexp)
      t = timed(lambda i: pagent.train(pexp), 3, lambda: None) / 3
      out['C5_pixel_ebm_train_step'] = {
          'ms_per_step': t * 1e3, 'images': b, 'images_per_s': b / t,
          'energy_evals_per_s': b * (n + 1) / t, 'tflops': flops / t * 1e-12,
          'precision': 'tf32', 'kernel_status': pnet.engine.kernel_status()}
    finally:
      gin.clear_config()
  # C5 Langevin variant (pushing_pixels/pixel_ebm_langevin.gin): 8 Langevin counter-examples
  # (100 iterations on the value head over one encoding), gradient penalty, Adam; and a
  # policy action (2048 samples, 100 + 100 "again" iterations, probabilities, arg max)
  if solo:
    from ibc_b200.networks.pixel_ebm import PixelEBM
    from ibc_b200.ibc.agents import ibc_policy
    gin.clear_config()
    gin.parse_config('DenseResnetValue.width = 1024\nDenseResnetValue.num_blocks = 1\n'
                     'langevin_actions_given_obs.late_fusion = True\n'
                     'langevin_actions_given_obs.num_iterations = 100')
    try:
      p_obs = {'rgb': specs.TensorSpec((2, 240, 320, 3), dtype=torch.uint8)}
      p_act = specs.BoundedTensorSpec((2,), minimum=-1.0, maximum=1.0)
      p_samp = specs.BoundedTensorSpec((2,), minimum=-1.1, maximum=1.1)
      lnet = PixelEBM((p_obs, p_act), specs.TensorSpec((1,)), 'ConvMaxpoolEncoder',
                      'DenseResnetValue', target_height=180, target_width=240, seed=0)
      ln, lk = 8, 100
      lagent = ibc_agent.ImplicitBCAgent(p_obs, p_act, p_samp, lnet, get_agent.Adam(1e-3),
                                         num_counter_examples=ln, fraction_langevin_samples=1.0,
                                         late_fusion=True, add_grad_penalty=True,
                                         run_full_chain_under_gradient=True, compute_mse=True)
      lexp = (({'rgb': img}, torch.rand(b, 2, device=device) * 2 - 1), ())
      for _ in range(2):
        lagent.train(lexp)
      t = timed(lambda i: lagent.train(lexp), 3, lambda: None) / 3
      out['C5_pixel_langevin_train_step'] = {
          'ms_per_step': t * 1e3, 'images': b, 'images_per_s': b / t, 'num_counter_examples': ln,
          'langevin_iterations': lk, 'grad_penalty': True,
          'energy_evals_per_s': (lk * b * ln + 2 * b * (ln + 1)) / t,
          'precision': 'tf32', 'kernel_status': lnet.engine.kernel_status()}
      lpol = ibc_policy.GreedyPolicy(ibc_policy.IbcPolicy(
          p_obs, p_act, p_samp, lnet, num_action_samples=2048, use_dfo=False, use_langevin=True,
          optimize_again=True, late_fusion=True))
      ts = specs.restart({'rgb': img[:1].contiguous()})
      for _ in range(2):
        lpol.action(ts)
      t = timed(lambda i: lpol.action(ts), 5, lambda: None) / 5
      out['C5_pixel_langevin_inference'] = {
          'ms_per_action': t * 1e3, 'actions_per_s': 1.0 / t, 'num_action_samples': 2048,
          'langevin_iterations': 2 * lk, 'kernel_status': lnet.engine.kernel_status()}
    finally:
      gin.clear_config()
  # C5 inference (pixel_ebm_best.gin: one frame pair, late fusion: encode once, DFO over 2048
  # candidates, 3 iterations) on the tf32 path
  pe.set_precision(_lib.PRECISION_TF32)
  mn, mx = torch.full((2,), -1.1, device=device), torch.full((2,), 1.1, device=device)
  frame = img[:1].contiguous()

  def pixel_action(i):
    enc = pe.encode(frame)
    cand = eng.uniform_actions(2048, mn, mx, None, 100 + i)
    probs = pe.dfo(enc, cand, 2048, mn, mx, 1.0, 3, 0.33, None, None, seed=200 + i)
    return cand[torch.argmax(probs)]
  for i in range(3):
    pixel_action(i)
  t = timed(pixel_action, 20, lambda: None) / 20
  out['C5_pixel_dfo_inference'] = {'ms_per_action': t * 1e3, 'actions_per_s': 1.0 / t,
                                   'num_action_samples': 2048, 'iterations': 3,
                                   'kernel_status': pe.kernel_status()}
  return out


def cpu_baseline(steps, warmup, batch=512):
  """Oracle (CPU restatement of the reference, PyTorch-CPU) train step on a
  bounded sample of the same workload: batch `batch` of 512, everything else
  identical.  kind = 'port' (the TF reference cannot be installed offline)."""
  from oracle import agent as oagent
  from oracle import mlp_ebm as omlp
  c = CFG
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  spec = omlp.MLPSpec(c['obs_dim'], c['act_dim'], c['width'], c['depth'])
  flat = omlp.init_params(spec, seed=1)
  state = oagent.adam_init(flat)
  cfg = oagent.AgentConfig(num_counter_examples=c['num_counter_examples'],
                           fraction_langevin_samples=0.0, add_grad_penalty=False)
  mn, mx = torch.full((c['act_dim'],), -1.1), torch.full((c['act_dim'],), 1.1)
  g = torch.Generator().manual_seed(0)
  n = c['num_counter_examples']

  def make_energy_fn(theta):
    def fn(obs_t, a):
      return omlp.energy(theta, spec, obs_t, a)
    return fn

  times = []
  for i in range(warmup + steps):
    obs = torch.randn(batch, c['obs_dim'], generator=g)
    act = torch.rand(batch, c['act_dim'], generator=g) * 2 - 1
    u = torch.rand(batch, n, c['act_dim'], generator=g)
    t0 = time.perf_counter()
    flat, loss, _, _ = oagent.train_step(make_energy_fn, flat, state, cfg, obs, act, mn, mx, u,
                                         c['lr'])
    times.append(time.perf_counter() - t0)
  t = float(np.mean(times[warmup:]))
  rows = batch * (n + 1)
  return {'value': rows / t, 'unit': 'energy evals/s', 'cores': cores, 'kind': 'port',
          'sample': 'oracle train step on batch %d of %d (N=%d, W=%d, D=%d), %d timed steps, '
                    '%.2f s/step' % (batch, c['batch'], n, c['width'], c['depth'], steps, t),
          'seconds_per_step': t}


def cpu_baselines_secondary():
  """the oracle (CPU restatement) beside C1, C3, C4 and C5 too, each
  on a bounded sample (reduced batch and, for the Langevin configs, reduced K: the cost of a
  chain is linear in K), all host cores, one warm-up + one timed step.  value = energy
  evaluations per second of the sample (SURVEY.md 8d count: B (N+1) for the DFO-style steps,
  K B N + 2 B (N+1) for the Langevin-style ones)."""
  from oracle import agent as oagent
  from oracle import mlp_ebm as omlp
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  out = {}
  cases = (
      ('C1_particle2d_train_step', omlp.MLPSpec(16, 2, 256, 2), 512, 256, 0, False, {}),
      ('C3_particle32_train_step', omlp.MLPSpec(256, 32, 256, 2), 512, 8, 50, True, {}),
      ('C4_d4rl_door_train_step', omlp.MLPSpec(39, 28, 512, 8), 128, 8, 20, True,
       dict(langevin_stepsize_init=0.5, langevin_delta_action_clip=0.5, langevin_noise_scale=0.5)),
  )
  for name, spec, b, n, k, gp, extra in cases:
    flat = omlp.init_params(spec, seed=1)
    state = oagent.adam_init(flat)
    cfg = oagent.AgentConfig(num_counter_examples=n, fraction_langevin_samples=1.0 if k else 0.0,
                             add_grad_penalty=gp, langevin_num_iterations=max(k, 1), **extra)
    mn, mx = torch.full((spec.act_dim,), -1.1), torch.full((spec.act_dim,), 1.1)
    g = torch.Generator().manual_seed(0)

    def make_energy_fn(theta, spec=spec):
      return lambda obs_t, a: omlp.energy(theta, spec, obs_t, a)

    t = None
    for it in range(2):
      obs = torch.randn(b, spec.obs_dim, generator=g)
      act = torch.rand(b, spec.act_dim, generator=g) * 2 - 1
      u = torch.rand(b, n, spec.act_dim, generator=g)
      draws = {'langevin_normals': torch.randn(k, b * n, spec.act_dim, generator=g)} if k else {}
      t0 = time.perf_counter()
      flat, _, _, _ = oagent.train_step(make_energy_fn, flat, state, cfg, obs, act, mn, mx, u, 1e-3,
                                        **draws)
      t = time.perf_counter() - t0
    evals = (k * b * n + 2 * b * (n + 1)) if k else b * (n + 1)
    out[name] = {'value': evals / t, 'unit': 'energy evals/s', 'cores': cores, 'kind': 'port',
                 'sample': 'oracle train step, batch %d of 512, N=%d, %s, %.2f s' % (
                     b, n, ('Langevin K=%d of 100 + gradient penalty' % k) if k else 'uniform negatives', t)}
  # C5: PixelEBM late-fusion forward + backward of the InfoNCE loss on 2 of 128 frame pairs
  try:
    from oracle import losses as olosses
    from oracle import pixel_ebm as opix
    spec = opix.PixelSpec() if hasattr(opix, 'PixelSpec') else None
    if spec is not None:
      flat = opix.init_params(spec, seed=1)
      b, n = 16, 256
      g = torch.Generator().manual_seed(0)
      img = torch.randint(0, 256, (b, spec.seq_len, spec.height, spec.width, spec.channels), generator=g,
                          dtype=torch.uint8)
      act = torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1
      t = None
      for it in range(2):
        theta = flat.clone().requires_grad_(True)
        t0 = time.perf_counter()
        e = opix.energy(theta, spec, img, act).reshape(b, n + 1)
        per = olosses.info_nce(e, b, n, 1.0)
        torch.autograd.grad(per.sum() / b, theta)
        t = time.perf_counter() - t0
      out['C5_pixel_ebm_train_grads'] = {
          'value': b * (n + 1) / t, 'images_per_s': b / t, 'unit': 'energy evals/s', 'cores': cores,
          'kind': 'port', 'sample': 'oracle PixelEBM loss + backward on %d of 128 frame pairs, N=256, %.2f s' % (b, t)}
  except Exception as ex:   # noqa: BLE001
    out['C5_pixel_ebm_train_grads'] = {'error': str(ex)[:160]}
  return out


def run_reference(args):
  """--impl reference: the reference's CPU implementation of the path.  The TF
  reference cannot be installed offline (no tensorflow / tf-agents wheels), so
  this times the oracle port on all host cores, on rank 0 only."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return None
  c = CFG
  # every one of the K timed steps runs on a bounded sample of the workload (a batch of b'
  # of the 512 observations, N and the network unchanged) sized from a calibration step so
  # that warm-up + K steps take about two minutes of host time
  cal = cpu_baseline(steps=1, warmup=1, batch=64)
  per_obs = cal['seconds_per_step'] / 64.0
  total = max(1, args.steps + min(args.warmup, 3))
  batch = int(max(8, min(c['batch'], 120.0 / (total * per_obs))))
  base = cpu_baseline(steps=max(1, args.steps), warmup=max(1, min(args.warmup, 3)), batch=batch)
  world = int(os.environ.get('WORLD_SIZE', '1'))
  return {
      'impl': 'reference', 'metric': 'ebm_train_energy_evals_per_s', 'value': base['value'],
      'unit': 'energy evals/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': base['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': workload_config(c),
      'cpu_baseline': base,
      'e2e': {'value': base['value'], 'unit': 'energy evals/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=500)   # ~0.5 s timed: enough nvidia-smi samples under load
  ap.add_argument('--warmup', type=int, default=20)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == 'b200':
    args.warmup = 3
  if args.impl == 'reference':
    res = run_reference(args)
  else:
    res = run_b200(args)
  if res is not None:
    print(json.dumps(res))


if __name__ == '__main__':
  main()
