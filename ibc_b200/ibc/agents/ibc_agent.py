"""ImplicitBCAgent on the B200 engine -- the reference's
`ibc/agents/ibc_agent.py:39-475` + `ibc/agents/base_agent.py:36-66` re-hosted.

`agent.train(experience)` runs: counter-example generation (uniform / DFO /
Langevin, `_make_counter_example_actions`), the B*(N+1)-row forward, InfoNCE
(+ gradient penalty), the backward pass, the data-parallel gradient all-reduce
(torch.distributed, NCCL on GPUs) and Adam -- all tensor math inside
libibc_b200.so.  experience = ((obs_dict [B, T, .], action [B, A]), ()).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch

from ibc_b200 import engine as eng
from ibc_b200 import gin_lite as gin
from ibc_b200 import parallel
from ibc_b200.ibc import specs
from ibc_b200.ibc.agents import ibc_policy
from ibc_b200.ibc.agents import mcmc
from ibc_b200.ibc.losses import ebm_loss
from ibc_b200.ibc.losses import gradient_loss
from ibc_b200.networks.mlp_ebm import MLPEBM
from ibc_b200.networks.pixel_ebm import PixelEBM


class BehavioralCloningAgent:
  """base_agent.py:29-66: loss -> gradients -> (all-reduce) -> optimizer."""

  def __init__(self, time_step_spec, action_spec, policy, collect_policy,
               train_step_counter=None):
    self.time_step_spec = time_step_spec
    self.action_spec = action_spec
    self.policy = policy
    self.collect_policy = collect_policy
    self.train_step_counter = 0 if train_step_counter is None else train_step_counter

  def train(self, experience, weights=None):
    return self._train(experience, weights)

  def get_eval_loss(self, experience):
    loss_dict = self._loss(experience, training=False)
    return loss_dict


_CONST_VECTORS = {}


def _scaled_sum(x, scale):
  """sum(x) * scale as one dot product with a cached constant vector."""
  key = (x.device, x.numel(), float(scale))
  c = _CONST_VECTORS.get(key)
  if c is None:
    if len(_CONST_VECTORS) > 64:
      _CONST_VECTORS.clear()
    c = torch.full((x.numel(),), float(scale), device=x.device, dtype=x.dtype)
    _CONST_VECTORS[key] = c
  return torch.dot(x.reshape(-1), c)


class _EarlyScalar(torch.Tensor):
  """A CUDA scalar whose VALUE can be read (float(), .item()) as soon as the kernels that produced
  it are done, without waiting for whatever the stream runs behind them: the value was also
  copied to pinned host memory on a second stream (TF eager reads loss_info.loss the same way:
  waiting for the op that produced it, not for the optimizer update enqueued after it).  Every
  other use is the plain device tensor."""
  __torch_function__ = torch._C._disabled_torch_function_impl

  @staticmethod
  def __new__(cls, dev, ring, slot, gen, ev):
    t = torch.Tensor._make_subclass(cls, dev)
    t._ring, t._slot, t._gen, t._ev = ring, slot, gen, ev
    return t

  def item(self):
    self._ev.synchronize()
    if self._ring.gen[self._slot] == self._gen:      # not yet overwritten by a later step
      return float(self._ring.host[self._slot])
    return self.as_subclass(torch.Tensor).item()

  def __float__(self):
    return float(self.item())

  def tolist(self):
    return self.item()


class _HostRing:
  """Pinned landing slots of the early loss copies (allocated once: cudaHostAlloc is slow)."""

  def __init__(self, n=64):
    self.host = torch.zeros(n, dtype=torch.float32, pin_memory=True)
    self.gen = [0] * n
    self.next = 0

  def take(self):
    s = self.next
    self.next = (s + 1) % len(self.gen)
    self.gen[s] += 1
    return s, self.gen[s]


def _scaled_sq_diff_sum(a, b, scale):
  d = (a - b).reshape(-1)
  return torch.dot(d, d) * scale


@gin.configurable
class ImplicitBCAgent(BehavioralCloningAgent):
  """TFAgent look-alike implementing implicit behavioral cloning."""

  def __init__(self,
               time_step_spec,
               action_spec,
               action_sampling_spec,
               cloning_network,
               optimizer,
               obs_norm_layer=None,
               act_norm_layer=None,
               act_denorm_layer=None,
               num_counter_examples=256,
               debug_summaries=False,
               summarize_grads_and_vars=False,
               train_step_counter=None,
               name=None,
               fraction_dfo_samples=0.,
               fraction_langevin_samples=1.0,
               ebm_loss_type='info_nce',
               late_fusion=False,
               compute_mse=False,
               run_full_chain_under_gradient=False,
               return_full_chain=True,
               add_grad_penalty=True,
               grad_norm_type='inf',
               softmax_temperature=1.0):
    del debug_summaries, summarize_grads_and_vars
    self.name = name
    self._action_sampling_spec = action_sampling_spec
    self._obs_norm_layer = obs_norm_layer
    self._act_norm_layer = act_norm_layer
    self._act_denorm_layer = act_denorm_layer
    self.cloning_network = cloning_network
    self.cloning_network.create_variables(training=False)
    self._optimizer = optimizer
    self._num_counter_examples = int(num_counter_examples)
    self._fraction_dfo_samples = fraction_dfo_samples
    self._fraction_langevin_samples = fraction_langevin_samples
    assert self._fraction_dfo_samples + self._fraction_langevin_samples <= 1.0
    assert self._fraction_dfo_samples >= 0.
    assert self._fraction_langevin_samples >= 0.
    self.ebm_loss_type = ebm_loss_type
    if ebm_loss_type == 'cd_kl':
      raise NotImplementedError("ebm_loss_type 'cd_kl' is out of scope (no shipped gin selects it; "
                                'it needs a copy of the network and gradients through the sampler)')
    if ebm_loss_type not in ('info_nce', 'cd', 'clipped_cd', 'soft_clipped_cd'):
      raise ValueError('Unsupported EBM loss type.')                 # ibc_agent.py:336
    if late_fusion != isinstance(cloning_network, PixelEBM):
      raise NotImplementedError('late_fusion must be used with (and only with) PixelEBM')
    self._run_full_chain_under_gradient = run_full_chain_under_gradient
    self._return_full_chain = return_full_chain
    self._add_grad_penalty = add_grad_penalty
    self._grad_norm_type = grad_norm_type
    self._softmax_temperature = softmax_temperature
    self._late_fusion = late_fusion
    self._compute_mse = compute_mse
    self._kl = None
    device = cloning_network.device
    self._min = torch.as_tensor(np.asarray(action_sampling_spec.minimum, np.float32)).to(device)
    self._max = torch.as_tensor(np.asarray(action_sampling_spec.maximum, np.float32)).to(device)
    collect_policy = ibc_policy.IbcPolicy(
        time_step_spec=time_step_spec,
        action_spec=action_spec,
        action_sampling_spec=action_sampling_spec,
        actor_network=cloning_network,
        late_fusion=late_fusion,
        obs_norm_layer=self._obs_norm_layer,
        act_denorm_layer=self._act_denorm_layer)
    policy = ibc_policy.GreedyPolicy(collect_policy)
    super().__init__(time_step_spec, action_spec, policy, collect_policy, train_step_counter)
    self._grads = torch.zeros_like(cloning_network.params)
    self.last_losses_dict = {}

  # -- helpers -----------------------------------------------------------------
  def _num_replicas(self) -> int:
    return parallel.num_replicas()

  def _loss_cfg(self):
    # grad_penalty's own gin bindings (gradient_loss.py:24-34)
    b = gin._BINDINGS.get('grad_penalty', {})
    if not b.get('only_apply_final_grad_penalty', True):
      raise NotImplementedError('only_apply_final_grad_penalty=False is not built for training')
    return eng.loss_cfg(self._softmax_temperature, self._add_grad_penalty, self._grad_norm_type,
                        b.get('grad_margin', 1.0), b.get('square_grad_penalty', True),
                        b.get('grad_loss_weight', 1.0), self.ebm_loss_type)

  def _to_device(self, experience):
    (observations, actions), _ = experience if len(experience) == 2 and isinstance(
        experience[0], tuple) else (experience, ())
    net = self.cloning_network
    obs = specs.map_structure(lambda t: t.to(net.device, torch.float32, non_blocking=True),
                              observations)
    act = actions.to(net.device, torch.float32, non_blocking=True)
    return obs, act

  # -- loss (ibc_agent.py:132-300) -------------------------------------------------
  def _loss(self, experience, variables_to_train=None, weights=None, training=False,
            draws: Optional[dict] = None):
    del variables_to_train
    if weights is not None:
      raise NotImplementedError('sample weights are not supported')
    net = self.cloning_network
    if self._late_fusion:
      return self._loss_late_fusion(experience, training, draws or {})
    observations, actions = self._to_device(experience)
    obs_flat = net._flat_obs(observations)                       # [B, O]
    batch_size = obs_flat.shape[0]
    n = self._num_counter_examples
    expanded_actions = actions[:, None, :]
    counter_example_actions, combined, chain_data = self._make_counter_example_actions(
        obs_flat, expanded_actions, batch_size, training, draws or {})
    losses_dict = {}
    global_batch = batch_size * self._num_replicas()            # aggregate_losses

    raw = eff = None
    if training and isinstance(net, MLPEBM):
      raw, eff = net.effective_params_for_training()             # spectral norm, training=True
    clipped = self.ebm_loss_type in ('clipped_cd', 'soft_clipped_cd')
    per_example, grad_loss, energies, grads = net.engine.train_grads(
        obs_flat, combined, n, self._loss_cfg(), global_batch, self._grads, want_energies=clipped)
    mse = None
    if self._compute_mse:
      # Keras MSE(NONE) -> [B, N]; aggregate_losses averages the extra axis then
      # sums / global batch (ibc_agent.py:182-189).  All groups have the same size,
      # so that is sum((c - a)^2) / (N * A * global_batch): two kernels instead of six.
      mse_scale = 1.0 / (float(global_batch) * counter_example_actions.shape[1]
                         * counter_example_actions.shape[2])
      mse = (lambda: _scaled_sq_diff_sum(counter_example_actions, expanded_actions, mse_scale),
             (counter_example_actions, expanded_actions))
    total_loss, mse_value = self._early_loss(net.engine, per_example, 1.0 / float(global_batch), mse)
    losses_dict['ebm_total_loss'] = total_loss
    if clipped:      # the logged scalars of clipped_cd (ibc_agent.py:252 losses_dict.update(debug_dict))
      losses_dict.update(ebm_loss.clipped_cd_debug(
          energies.view(batch_size, n + 1), counter_example_actions,
          expanded_actions.expand(batch_size, n, expanded_actions.shape[-1]),
          self.ebm_loss_type == 'soft_clipped_cd'))
    if self._add_grad_penalty:
      losses_dict['grad_loss'] = grad_loss.mean()
    if self._compute_mse:
      losses_dict['mse_counter_examples'] = mse_value
    self._chain_metrics(chain_data, losses_dict)
    self.last_losses_dict = losses_dict
    if training:
      if raw is not None:   # chain dL/dW_bar through the spectral normalisation
        (graw,) = torch.autograd.grad(eff, raw, grad_outputs=grads)
        grads.copy_(graw)
      return specs.LossInfo(total_loss, ()), grads
    return losses_dict

  def _early_loss(self, engine, per_example, scale, extra=None):
    """sum(per_example) * scale.  When the engine recorded its loss event (the per-example losses
    are final while the backward is still running) the sum and a copy to pinned host memory run
    on a second stream behind that event, and the returned tensor's float() waits for that copy
    only: a caller that reads the loss every step no longer drains the whole step before it can
    submit the next one.
    extra = (fn, inputs): a second logged scalar fn() of tensors that exist before the forward
    (the counter examples' MSE), computed in the same place -- beside the backward instead of as
    small launches between the backward and Adam.  Returns (loss, extra value or None)."""
    if not per_example.is_cuda or not hasattr(engine, 'stream_wait_loss'):
      return _scaled_sum(per_example, scale), (extra[0]() if extra else None)
    from ibc_b200 import parallel
    side = parallel._side_stream(per_example.device)
    if not engine.stream_wait_loss(side):
      return _scaled_sum(per_example, scale), (extra[0]() if extra else None)
    ring = getattr(self, '_loss_ring', None)
    if ring is None:
      ring = self._loss_ring = _HostRing()
    slot, gen = ring.take()
    slot2, gen2 = ring.take() if extra else (None, None)
    main = torch.cuda.current_stream(per_example.device)
    val2 = None
    with torch.cuda.stream(side):
      loss = _scaled_sum(per_example, scale)
      ring.host[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)
      # the loss's event FIRST: a caller that reads the loss every step is released before the
      # extra scalar is computed (one event behind both: end-to-end step 0.485 -> 0.528 ms)
      ev = ev2 = torch.cuda.Event()
      ev.record(side)
      if extra:
        val2 = extra[0]()
        ring.host[slot2:slot2 + 1].copy_(val2.reshape(1), non_blocking=True)
        ev2 = torch.cuda.Event()
        ev2.record(side)
    per_example.record_stream(side)
    loss.record_stream(main)
    if extra:
      for t in extra[1]:
        t.record_stream(side)
      val2.record_stream(main)
    main.wait_stream(side)             # later uses of the device values on the main stream
    return (_EarlyScalar(loss, ring, slot, gen, ev),
            _EarlyScalar(val2, ring, slot2, gen2, ev2) if extra else None)

  @staticmethod
  def _chain_metrics(chain_data, losses_dict):
    """ibc_agent.py:263-280: logged averages over the training-time chain."""
    if chain_data is not None and chain_data.energies is not None:
      e = chain_data.energies
      losses_dict['overall_energies_avg'] = e.mean()
      losses_dict['first_energies_avg'] = e[0].mean()
      losses_dict['final_energies_avg'] = e[-1].mean()
    if chain_data is not None and chain_data.grad_norms is not None:
      g = chain_data.grad_norms
      losses_dict['overall_grad_norms_avg'] = g.mean()
      losses_dict['first_grad_norms_avg'] = g[0].mean()
      losses_dict['final_grad_norms_avg'] = g[-1].mean()

  def _loss_late_fusion(self, experience, training, draws):
    """ibc_agent.py:191-213: encode once (training=True), tile the encoding N+1
    times, value head, InfoNCE; gradients flow into the conv encoder."""
    net = self.cloning_network
    (observations, actions), _ = experience if len(experience) == 2 and isinstance(
        experience[0], tuple) else (experience, ())
    images = net._images(observations)
    actions = actions.to(net.device, torch.float32, non_blocking=True)
    batch_size = images.shape[0]
    n = self._num_counter_examples
    expanded_actions = actions[:, None, :]
    # negatives: uniform, optionally refined by DFO / Langevin on the value head over one
    # encoding of the un-tiled observations (ibc_agent.py:340-475 with late fusion)
    global_batch = batch_size * self._num_replicas()
    if self._fraction_dfo_samples > 0. or self._fraction_langevin_samples > 0.:
      # one encoder pass serves the sampler and the backward (upstream encodes twice, with
      # identical results: the encoder has no dropout / batch norm)
      self._enc = net.engine.train_begin(images, n)
      try:
        counter, combined, chain_data = self._make_counter_example_actions(
            images, expanded_actions, batch_size, training, draws)
      finally:
        self._enc = None
      per_example, grad_loss, _, grads = net.engine.train_finish(
          combined, self._loss_cfg(), global_batch, self._grads)
    else:
      counter, combined, chain_data = self._make_counter_example_actions(
          images, expanded_actions, batch_size, training, draws)
      per_example, grad_loss, _, grads = net.engine.train_grads(
          images, combined, n, self._loss_cfg(), global_batch, self._grads, want_grad_loss=True)
    total_loss = _scaled_sum(per_example, 1.0 / float(global_batch))
    losses_dict = {'ebm_total_loss': total_loss}
    if self._add_grad_penalty:
      losses_dict['grad_loss'] = grad_loss.mean()
    self._chain_metrics(chain_data, losses_dict)
    if self._compute_mse:
      losses_dict['mse_counter_examples'] = _scaled_sq_diff_sum(
          counter, expanded_actions,
          1.0 / (float(global_batch) * counter.shape[1] * counter.shape[2]))
    self.last_losses_dict = losses_dict
    if training:
      return specs.LossInfo(total_loss, ()), grads
    return losses_dict

  def _compute_ebm_loss(self, batch_size, predictions, counter_example_actions,
                        expanded_actions, grad_flow_network_inputs=None):
    """ibc_agent.py:302-338 (forward value only)."""
    del grad_flow_network_inputs
    if self.ebm_loss_type == 'info_nce':
      return ebm_loss.info_nce(predictions, batch_size, self._num_counter_examples,
                               self._softmax_temperature, self._kl)
    if self.ebm_loss_type == 'cd':
      return ebm_loss.cd(predictions)
    if self.ebm_loss_type in ('clipped_cd', 'soft_clipped_cd'):
      actions_size_n = expanded_actions.expand(batch_size, self._num_counter_examples,
                                               expanded_actions.shape[-1])
      return ebm_loss.clipped_cd(predictions, counter_example_actions, actions_size_n,
                                 soft=self.ebm_loss_type == 'soft_clipped_cd')
    raise ValueError('Unsupported EBM loss type.')

  # -- negatives (ibc_agent.py:340-475) -----------------------------------------------
  def _make_counter_example_actions(self, obs_flat, expanded_actions, batch_size,
                                    training=False, draws=None):
    """obs_flat [B, O] (un-tiled; the frames [B, T, H, W, C] with late fusion),
    expanded_actions [B, 1, A].  Returns
    (counter_example_actions [B, N, A], combined [B*(N+1), A], chain_data)."""
    del training   # sampler network calls always use training=False (:388, :403)
    draws = draws or {}
    net = self.cloning_network
    n = self._num_counter_examples
    a_dim = expanded_actions.shape[-1]
    u = draws.get('uniform_u')
    if u is not None:
      u = u.reshape(batch_size * n, a_dim).contiguous()
    chain_data = None
    if self._fraction_dfo_samples == 0.0 and self._fraction_langevin_samples == 0.0:
      # uniform negatives only: samples and the concatenation with the expert action in one
      # launch (the same samples as uniform_actions; two torch.cat launches less in front of
      # the step's first large kernel)
      combined = eng.counter_examples_uniform(
          expanded_actions.reshape(batch_size, a_dim).contiguous(), n, self._min, self._max, u,
          mcmc._next_seed())
      counter = combined.view(batch_size, n + 1, a_dim)[:, :n]
      return counter, combined, chain_data
    random_uniform = eng.uniform_actions(batch_size * n, self._min, self._max, u, mcmc._next_seed())
    dfo_actions = lang_actions = None
    if self._fraction_dfo_samples > 0.:
      if self._return_full_chain:
        raise NotImplementedError('Not implemented to return dfo chain.')   # :374-375
      dfo_actions = random_uniform.clone()
      dfo_obs = obs_flat
      if self._late_fusion:
        dfo_obs = getattr(self, '_enc', None)
        if dfo_obs is None:
          dfo_obs = net.encode(obs_flat, training=False)
      net.engine.dfo(dfo_obs, dfo_actions, n, self._min, self._max, 1.0, 3, 0.33,
                     draws.get('dfo_uniforms'), draws.get('dfo_normals'), mcmc._next_seed())
    if self._fraction_langevin_samples > 0.:
      # late fusion: un-tiled frames, encoded once inside the sampler (mcmc.py:384-388;
      # upstream binds langevin_actions_given_obs.late_fusion in the gin file)
      # the engine samplers take un-tiled observations (mcmc._untile passes them through);
      # only a generic callable network needs the B*n-row tile_batch copy
      tiled = obs_flat if (self._late_fusion or isinstance(net, MLPEBM)) \
          else torch.repeat_interleave(obs_flat, n, dim=0)
      lf = {'late_fusion': True,
            'observation_encoding': getattr(self, '_enc', None)} if self._late_fusion else {}
      out = mcmc.langevin_actions_given_obs(
          net, tiled, random_uniform, policy_state=(), min_actions=self._min,
          max_actions=self._max, num_action_samples=n, training=False,
          return_chain=self._return_full_chain, grad_norm_type=self._grad_norm_type,
          normal_draws=draws.get('langevin_normals'), **lf)
      if self._return_full_chain:
        lang_actions, chain_data = out
      else:
        lang_actions = out
    frac_init = 1.0 - self._fraction_dfo_samples - self._fraction_langevin_samples
    n_init = int(frac_init * n)
    n_dfo = int(self._fraction_dfo_samples * n)
    n_lang = int(self._fraction_langevin_samples * n)
    n_init += n - n_init - n_dfo - n_lang                    # residual -> init (:416-430)
    pieces, used = [], 0
    if n_init > 0:
      pieces.append(random_uniform.view(batch_size, n, a_dim)[:, :n_init])
      used += n_init
    if n_dfo > 0:
      pieces.append(dfo_actions.view(batch_size, n, a_dim)[:, used:used + n_dfo])
      used += n_dfo
    if n_lang > 0:
      pieces.append(lang_actions.view(batch_size, n, a_dim)[:, used:used + n_lang])
      used += n_lang
    counter = torch.cat(pieces, dim=1)
    combined = torch.cat([counter, expanded_actions], dim=1)    # true action LAST (:461-473)
    combined = combined.reshape(batch_size * (n + 1), a_dim).contiguous()
    return counter, combined, chain_data

  # -- train (base_agent.py:36-62) ------------------------------------------------------
  # check_numerics (base_agent.py:46) is an in-graph op upstream: it does not stall
  # the Python thread.  'deferred' keeps that property (the loss of step k is
  # checked, without a device sync of its own, when step k+1 is submitted or when
  # the caller reads it); 'sync' raises inside the offending step.
  check_numerics = 'deferred'

  def _check_finite(self, loss, extra_host=None):
    """extra_host: a pinned one-element tensor a kernel of this step writes (the peer all-reduce's
    reduced d/d be): looked at together with the loss, behind an event recorded now."""
    if self.check_numerics == 'sync':
      if not math.isfinite(float(loss)):
        raise FloatingPointError('Loss is inf or nan')
      if extra_host is not None:
        torch.cuda.current_stream().synchronize()
        if not math.isfinite(float(extra_host[0])):
          raise FloatingPointError('Loss is inf or nan')
      return
    # Deferred: the loss is copied to pinned host memory behind the step's kernels and
    # looked at (a plain host read, no CUDA call, no stream sync) once its event has
    # completed -- typically while the next step is being submitted.
    pending = getattr(self, '_pending_finite', None)
    if pending is not None:
      hosts, ev = pending
      if not ev.query():
        return                        # still in flight: keep waiting on the older one
      self._pending_finite = None
      if not all(math.isfinite(float(h[0])) for h in hosts):
        raise FloatingPointError('Loss is inf or nan')
    if isinstance(loss, _EarlyScalar) and loss._ring.gen[loss._slot] == loss._gen:
      # the early read-out already copies this loss to pinned memory: look at that copy
      hosts, ev = [loss._ring.host[loss._slot:loss._slot + 1]], loss._ev
      if extra_host is not None:
        hosts.append(extra_host)
        ev = torch.cuda.Event()
        ev.record()
      self._pending_finite = (hosts, ev)
      return
    host = getattr(self, '_finite_host', None)
    if host is None:
      host = torch.empty(1, dtype=torch.float32, pin_memory=True)
      self._finite_host = host
    host.copy_(loss.detach().reshape(1), non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    self._pending_finite = ([host] + ([extra_host] if extra_host is not None else []), ev)

  def flush_checks(self):
    """Look at the last deferred finiteness check (call when training ends)."""
    pending = getattr(self, '_pending_finite', None)
    if pending is not None:
      hosts, ev = pending
      ev.synchronize()
      self._pending_finite = None
      if not all(math.isfinite(float(h[0])) for h in hosts):
        raise FloatingPointError('Loss is inf or nan')

  def _peer_reducer(self):
    """parallel.PeerReducer of this agent, created on the first multi-replica step (None: NCCL).
    Only where the flat gradient leaves the engine as it is all-reduced (no spectral-norm chain
    rule in between) and Adam is the plain optimizer class."""
    if not hasattr(self, '_peer'):
      net = self.cloning_network
      plain = (parallel.num_replicas() > 1 and net.params.is_cuda
               and getattr(net, '_sn_u', None) is None
               and hasattr(self._optimizer, 'apply_gradients_peer'))
      self._peer = parallel.PeerReducer.create(net.params.numel(), net.params.device) if plain else None
    return self._peer

  def _train(self, experience, weights=None, draws=None):
    peer = self._peer_reducer()
    if peer is not None:
      # gradients go straight into this step's peer-visible buffer; sum + Adam are one launch
      step = self._optimizer.iterations
      self._grads = peer.buffer(step)
      loss_info, grads = self._loss(experience, weights=weights, training=True, draws=draws)
      self._optimizer.apply_gradients_peer(self.cloning_network, peer, step)
      # d/d be summed over ALL replicas is non-finite as soon as any replica's loss is: every
      # rank sees it in the reduced last gradient entry, which the kernel writes to pinned host
      # memory (no probe kernels between this step and the next)
      self._check_finite(loss_info.loss, extra_host=peer.tail)
      self.train_step_counter += 1
      return loss_info
    loss_info, grads = self._loss(experience, weights=weights, training=True, draws=draws)
    parallel.allreduce_gradients(grads)                         # MirroredStrategy all-reduce
    probe = loss_info.loss
    if parallel.num_replicas() > 1:
      # every rank must raise on the same step (a rank that raised before the all-reduce would
      # leave the others blocked in NCCL): d/d be = sum of dE over ALL replicas is non-finite
      # as soon as any replica's loss is, and 0 * (inf | nan) = nan
      probe = probe + grads[-1] * 0.0
    self._check_finite(probe)
    self._optimizer.apply_gradients(self.cloning_network, grads)
    self.train_step_counter += 1
    return loss_info
