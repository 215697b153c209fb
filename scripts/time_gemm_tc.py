"""Diagnostic: throughput of the general tcgen05 tf32 GEMM (csrc/gemm_tc.cu) on the
contractions of the C5 PixelEBM step (im2col rows of the four convolutions at batch 16,
the value head at 32 896 rows) in their three forms (forward NN, data gradient NT,
weight gradient TN with split-K atomics), and the PixelEBM train step in both
precisions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ibc_b200 import _lib
from ibc_b200 import engine as eng


def timed(fn, reps=10, warm=3):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(reps):
    fn()
  t.record(); torch.cuda.synchronize()
  return s.elapsed_time(t) / reps


nb = 16
shapes = [('conv1 rows', nb * 180 * 240, 32, 54), ('conv2 rows', nb * 90 * 120, 64, 288),
          ('conv3 rows', nb * 45 * 60, 128, 576), ('conv4 rows', nb * 22 * 30, 256, 1152),
          ('value d0', 32896, 256, 1024), ('value d2', 32896, 1024, 256),
          ('dense 4096', 4096, 4096, 4096)]
for name, m, n, k in shapes:
  a = torch.randn(m, k, device='cuda')
  w = torch.randn(k, n, device='cuda')
  g = torch.randn(m, n, device='cuda')
  fl = 2.0 * m * n * k
  t_nn = timed(lambda: eng.gemm_tf32(a, w, m, n, k, check=False))
  t_nt = timed(lambda: eng.gemm_tf32(g, w, m, k, n, trans_b=True, check=False))
  dw = torch.zeros(k, n, device='cuda')
  t_tn = timed(lambda: eng.gemm_tf32(a, g, k, n, m, trans_a=True, out=dw, check=False))
  print('%-12s M=%7d N=%5d K=%5d  NN %.3f ms %6.1f TF/s | NT %.3f ms %6.1f TF/s | TN %.3f ms %6.1f TF/s'
        % (name, m, n, k, t_nn, fl / t_nn * 1e-9, t_nt, fl / t_nt * 1e-9, t_tn, fl / t_tn * 1e-9))

# C5 train step (pixel_ebm_best.gin shapes) at a few batch sizes, both precisions
for b in (32, 128):
  e = eng.PixelEngine(2, 240, 320, 3, 180, 240, 2, 1024, 1)
  flat = (torch.randn(e.param_count) * 0.05).cuda()
  e.bind_params(flat)
  img = torch.randint(0, 256, (b, 2, 240, 320, 3), dtype=torch.uint8).cuda()
  act = (torch.rand(b * 257, 2) * 2 - 1).cuda()
  cfg = eng.loss_cfg(1.0, False)
  for prec, label in ((_lib.PRECISION_TF32, 'tf32 tcgen05'), (_lib.PRECISION_FP32, 'fp32 FMA')):
    e.set_precision(prec)
    ms = timed(lambda: e.train_grads(img, act, 256, cfg, b), reps=3, warm=1)
    t_enc = timed(lambda: e.encode(img), reps=3, warm=1)
    print('C5 PixelEBM B=%3d train_grads %.2f ms (%.0f images/s), encode %.2f ms  [%s] status %d'
          % (b, ms, b / ms * 1e3, t_enc, label, e.kernel_status()))
  del e
