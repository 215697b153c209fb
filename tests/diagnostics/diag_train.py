"""Diagnostic: per-tensor relative error of train_grads (tf32 and fp32) vs oracle autograd."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import mlp_ebm as omlp, losses as olosses
from ibc_b200 import engine as eng

def run(spec, b, n, scale=1.5):
  flat = omlp.init_params(spec, seed=1) * scale
  g = torch.Generator().manual_seed(23)
  obs = torch.randn(b, spec.obs_dim, generator=g)
  act = torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1
  theta = flat.double().clone().requires_grad_(True)
  obs_t = torch.repeat_interleave(obs, n + 1, 0).double()
  pred = omlp.energy(theta, spec, obs_t, act.double()).reshape(b, n + 1)
  per = olosses.info_nce(pred, b, n, 1.0)
  (want,) = torch.autograd.grad(per.sum() / b, theta)
  grms = float(want.pow(2).mean().sqrt())
  for prec in (1, 0):
    e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
    dev = flat.cuda().contiguous(); e.bind_params(dev); e.set_precision(prec)
    per_g, gl, en, grads = e.train_grads(obs.cuda(), act.cuda(), n, eng.loss_cfg(1.0, False), b, want_energies=True)
    st = e.kernel_status()
    out = []
    for name, off, shape in omlp.param_layout(spec):
      cnt = math.prod(shape)
      w = want[off:off + cnt]; gg = grads[off:off + cnt].double().cpu()
      rel = float((gg - w).norm()) / (float(w.norm()) + 1e-30)
      relf = float((gg - w).norm()) / (float(w.norm()) + 1e-2 * grms * math.sqrt(cnt))
      out.append('%s:%.2g/%.2g' % (name, rel, relf))
    print(spec, (b, n), 'prec', prec, 'status', st, ' '.join(out), flush=True)
    if spec.act_dim > 2 and prec == 1:
      o = spec.obs_dim * spec.width
      w = want[:o]; gg = grads[:o].double().cpu()
      print('   Wp obs part rel %.3g ; act part rel %.3g' % (float((gg - w).norm() / w.norm()),
            float((grads[o:(spec.obs_dim + spec.act_dim) * spec.width].double().cpu() - want[o:(spec.obs_dim + spec.act_dim) * spec.width]).norm() / want[o:(spec.obs_dim + spec.act_dim) * spec.width].norm())))

run(omlp.MLPSpec(20, 2, 128, 2), 8, 7)
run(omlp.MLPSpec(39, 28, 128, 4), 5, 100)
run(omlp.MLPSpec(39, 28, 128, 4), 8, 31)
run(omlp.MLPSpec(20, 2, 128, 16), 16, 63)
