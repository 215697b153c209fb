"""How far apart do two correct implementations of the C4 Langevin chain end up?  The oracle in
fp64 against the oracle in fp32 and against its tf32-operand emulation, next to the GPU paths
(K = 100, supplied normals).  Calibrates tests/test_gpu_parity.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import engine as eng
from oracle import mcmc as omcmc, mlp_ebm as omlp, tf32_emul

torch.set_num_threads(os.cpu_count() or 1)
torch.cuda.set_device(0)
for name, spec, kw in (('C4', omlp.MLPSpec(39, 28, 512, 8), dict(noise_scale=0.5, delta_action_clip=0.5, sampler_stepsize_init=0.5)),
                       ('C3', omlp.MLPSpec(256, 32, 256, 2), dict())):
  flat = omlp.init_params(spec, seed=1)
  b, n, k = 128, 8, 100
  g = torch.Generator().manual_seed(13)
  obs = torch.randn(b, spec.obs_dim, generator=g); act0 = torch.rand(b * n, spec.act_dim, generator=g) * 2 - 1
  mn, mx = torch.full((spec.act_dim,), -1.1), torch.full((spec.act_dim,), 1.1)
  normals = torch.randn(k, b * n, spec.act_dim, generator=torch.Generator().manual_seed(3))
  obs_t = torch.repeat_interleave(obs, n, 0)
  res = {}
  for label, fn, dt in (('fp64', omlp.energy, torch.float64), ('fp32', omlp.energy, torch.float32),
                        ('tf32emul', tf32_emul.energy, torch.float64)):
    t0 = time.time()
    f = flat.to(dt); o = obs_t.to(dt)
    res[label] = omcmc.langevin_actions_given_obs(lambda _o, a: fn(f, spec, o, a), None, act0.to(dt), mn.to(dt), mx.to(dt),
                                                  normals, num_iterations=k, **kw).double()
    print(name, label, 'oracle %.1f s' % (time.time() - t0), flush=True)
  for prec, label in ((0, 'gpu_fp32'), (1, 'gpu_tf32')):
    e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
    dev = flat.cuda().contiguous(); e.bind_params(dev); e.set_precision(prec)
    a = act0.cuda().clone()
    e.langevin(obs.cuda(), a, n, mn.cuda(), mx.cuda(), eng.langevin_cfg(num_iterations=k, **kw), normals.cuda())
    res[label] = a.double().cpu()
  ref = res['fp64']
  for label in ('fp32', 'tf32emul', 'gpu_fp32', 'gpu_tf32'):
    d = (res[label] - ref).abs().amax(dim=1)
    print('%s %-9s vs fp64: mean %.2e median %.2e p99 %.2e max %.2e frac>1.1e-2 %.4f' % (
        name, label, d.mean(), d.median(), d.quantile(0.99), d.max(), (d > 1.1e-2).float().mean()))
  d = (res['gpu_tf32'] - res['tf32emul']).abs().amax(dim=1)
  print('%s gpu_tf32 vs tf32emul: mean %.2e max %.2e' % (name, d.mean(), d.max()))
