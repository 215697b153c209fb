"""Diagnostic (not a test): tf32 energy error statistics per shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import mlp_ebm as omlp
from ibc_b200 import engine as eng, _lib

def run(spec, b, n, scale=2.0):
  e = eng.EBMEngine(spec.obs_dim, spec.act_dim, spec.width, spec.depth)
  flat = omlp.init_params(spec, seed=1) * scale
  dev = flat.cuda().contiguous(); e.bind_params(dev)
  g = torch.Generator().manual_seed(0)
  obs = torch.randn(b, spec.obs_dim, generator=g)
  act = torch.rand(b * n, spec.act_dim, generator=g) * 2 - 1
  want = omlp.energy(flat.double(), spec, torch.repeat_interleave(obs, n, 0).double(), act.double())
  out = {}
  for prec in (1, 0):
    e.set_precision(prec)
    got = e.energy(obs.cuda(), act.cuda(), n).double().cpu()
    st = e.kernel_status()
    err = (got - want).abs()
    rms = float(want.pow(2).mean().sqrt())
    ratio = err / (want.abs() + rms)
    bad = (ratio > 1e-2).nonzero().reshape(-1)
    out[prec] = (float(ratio.max()), float(ratio.mean()), st, bad[:8].tolist(), int(bad.numel()))
  print(spec, 'b,n=', (b, n), 'scale', scale, 'rms=%.3g' % rms, 'tf32 max/mean ratio %.3g %.3g status %d bad %s (#%d) | fp32 max %.3g' % (
      out[1][0], out[1][1], out[1][2], out[1][3], out[1][4], out[0][0]), flush=True)

for spec in [omlp.MLPSpec(20, 2, 128, 2), omlp.MLPSpec(20, 2, 128, 16), omlp.MLPSpec(16, 2, 256, 2),
             omlp.MLPSpec(256, 32, 256, 2)]:
  for (b, n) in [(3, 5), (2, 256), (5, 300), (1, 2048), (512, 257)]:
    for scale in (1.0, 2.0):
      run(spec, b, n, scale)
