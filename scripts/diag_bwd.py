"""Fused (block-major, fp16 weight-gradient GEMM) and split (tf32) backward against the
fp64 oracle and its tf32 emulation: relative Frobenius error per parameter tensor."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import _lib, engine as eng
from oracle import losses as olosses, mlp_ebm as omlp, tf32_emul

torch.cuda.set_device(0)
for spec, b, n, scale in ((omlp.MLPSpec(20, 2, 128, 16), 3, 50, 1.0), (omlp.MLPSpec(39, 28, 128, 2), 37, 100, 1.5),
                          (omlp.MLPSpec(20, 2, 128, 4), 40, 256, 1.5)):
  flat = omlp.init_params(spec, seed=1) * scale
  e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
  dev = flat.cuda().contiguous(); e.bind_params(dev); e.set_precision(_lib.PRECISION_TF32)
  g = torch.Generator().manual_seed(29)
  obs = torch.randn(b, spec.obs_dim, generator=g); act = torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1
  obs_t = torch.repeat_interleave(obs, n + 1, 0).double()
  refs = []
  for fn in (omlp.energy, tf32_emul.energy):
    theta = flat.double().clone().requires_grad_(True)
    pred = fn(theta, spec, obs_t, act.double()).reshape(b, n + 1)
    per = olosses.info_nce(pred, b, n, 1.0)
    (gr,) = torch.autograd.grad(per.sum() / b, theta)
    refs.append(gr)
  cfg = eng.loss_cfg(1.0, False)
  out = {}
  for mode in ('fused', 'split'):
    if mode == 'split': os.environ['IBC_TRAIN_SPLIT'] = '1'
    else: os.environ.pop('IBC_TRAIN_SPLIT', None)
    _, _, _, gg = e.train_grads(obs.cuda(), act.cuda(), n, cfg, b)
    out[mode] = gg.double().cpu().clone()
    print(mode, 'status', e.kernel_status())
  grms = float(refs[0].pow(2).mean().sqrt())
  print(spec, b, n)
  for name, off, shape in omlp.param_layout(spec):
    cnt = math.prod(shape)
    w = refs[0][off:off + cnt]
    den = float(w.norm()) + 1e-2 * grms * math.sqrt(cnt)
    r = lambda x: float((x[off:off + cnt] - w).norm()) / den
    print('  %-6s emul %.2e  fused %.2e  split %.2e  fused-split %.2e' % (
        name, r(refs[1]), r(out['fused']), r(out['split']),
        float((out['fused'][off:off + cnt] - out['split'][off:off + cnt]).norm()) / den))

# full C2 size: both against the exact fp32 FMA path of the same handle
for spec, b, n, scale in ((omlp.MLPSpec(20, 2, 128, 16), 512, 256, 1.0), (omlp.MLPSpec(20, 2, 128, 4), 160, 256, 1.5)):
  flat = omlp.init_params(spec, seed=1) * scale
  e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
  dev = flat.cuda().contiguous(); e.bind_params(dev)
  g = torch.Generator().manual_seed(29)
  obs = torch.randn(b, spec.obs_dim, generator=g).cuda(); act = (torch.rand(b * (n + 1), spec.act_dim, generator=g) * 2 - 1).cuda()
  cfg = eng.loss_cfg(1.0, False)
  e.set_precision(_lib.PRECISION_FP32)
  _, _, _, gg = e.train_grads(obs, act, n, cfg, b)
  ref = gg.double().cpu().clone()
  e.set_precision(_lib.PRECISION_TF32)
  out = {}
  for mode in ('fused', 'split'):
    if mode == 'split': os.environ['IBC_TRAIN_SPLIT'] = '1'
    else: os.environ.pop('IBC_TRAIN_SPLIT', None)
    _, _, _, gg = e.train_grads(obs, act, n, cfg, b)
    out[mode] = gg.double().cpu().clone()
    print(mode, 'status', e.kernel_status())
  grms = float(ref.pow(2).mean().sqrt())
  print(spec, b, n, '(reference: fp32 path)')
  for name, off, shape in omlp.param_layout(spec):
    cnt = math.prod(shape)
    w = ref[off:off + cnt]
    den = float(w.norm()) + 1e-2 * grms * math.sqrt(cnt)
    r = lambda x: float((x[off:off + cnt] - w).norm()) / den
    print('  %-6s fused %.2e  split %.2e  fused-split %.2e' % (
        name, r(out['fused']), r(out['split']),
        float((out['fused'][off:off + cnt] - out['split'][off:off + cnt]).norm()) / den))
