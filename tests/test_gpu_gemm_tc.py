"""GPU parity tests of the general tcgen05 tf32 GEMM (ibc_gemm_tf32, csrc/gemm_tc.cu): the
Dense / Conv2D contraction of networks/layers/dense_resnet_value.py:31-69 and
networks/layers/conv_maxpool.py:20-48 with its fused prologue / epilogue.

This is a floating-point kernel, so the reference is plain PyTorch: the operands rounded
to nearest tf32 exactly as the kernel rounds them (oracle/tf32_emul.round_tf32), the
product in fp64.  What is left is fp32 accumulation order inside the tensor core:
|d| <= 5e-5 * (|x| + rms(x)) (measured worst 2.2e-5 at K = 4096).  Against the un-rounded
fp64 product the bound is the tf32 operand error, 2e-3 of the same form."""
import pytest
import torch

from oracle.tf32_emul import round_tf32

pytestmark = pytest.mark.gpu

EXACT_TOL = 5e-5
TF32_TOL = 2e-3


def _close(got, want, tol, name=''):
  got, want = got.detach().double().cpu(), want.detach().double().cpu()
  assert got.shape == want.shape, (name, got.shape, want.shape)
  rms = float(want.pow(2).mean().sqrt())
  worst = float(((got - want).abs() / (tol * (want.abs() + rms) + 1e-30)).max())
  assert worst <= 1.0, '%s: err/bound %.3g (rms %.3g)' % (name, worst, rms)


def _operands(m, n, k, ta, tb, seed, pad_a=0, pad_b=0):
  """Row-major operands whose leading dimensions carry `pad` extra columns."""
  g = torch.Generator().manual_seed(seed)
  a_shape = (k, m) if ta else (m, k)
  b_shape = (n, k) if tb else (k, n)
  a_full = torch.randn(a_shape[0], a_shape[1] + pad_a, generator=g)
  b_full = torch.randn(b_shape[0], b_shape[1] + pad_b, generator=g)
  a, b = a_full[:, :a_shape[1]], b_full[:, :b_shape[1]]
  return a_full.cuda()[:, :a_shape[1]], b_full.cuda()[:, :b_shape[1]], a, b


def _ref(a, b, ta, tb, relu_a=False, rounded=True):
  a = a.t() if ta else a
  b = b.t() if tb else b
  if relu_a:
    a = a.clamp_min(0)
  if rounded:
    a, b = round_tf32(a.contiguous()), round_tf32(b.contiguous())
  return a.double() @ b.double()


SHAPES = [
    (128, 128, 32), (128, 128, 256), (256, 384, 160), (1000, 72, 54), (131, 32, 96),
    (64, 1024, 258), (77, 19, 33), (4096, 256, 1024), (300, 64, 288), (129, 257, 4096),
]


@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize('m,n,k', SHAPES)
def test_plain_product_all_layouts(m, n, k, ta, tb):
  from ibc_b200 import engine as eng
  ag, bg, a, b = _operands(m, n, k, ta, tb, seed=m + 7 * n + 13 * k)
  c, status = eng.gemm_tf32(ag, bg, m, n, k, trans_a=ta, trans_b=tb)
  assert status == 0, 'barrier timeout code %d' % status
  _close(c, _ref(a, b, ta, tb), EXACT_TOL, 'rounded-operand product')
  _close(c, _ref(a, b, ta, tb, rounded=False), TF32_TOL, 'exact product')


@pytest.mark.parametrize('ta,tb', [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize('pad_a,pad_b', [(1, 0), (0, 3), (2, 2), (4, 8)])
def test_unaligned_leading_dimensions_take_the_scalar_path(ta, tb, pad_a, pad_b):
  """Leading dimensions that are not multiples of 4 floats (im2col rows of the first
  convolution: K = 54) cannot use 16-byte loads."""
  from ibc_b200 import engine as eng
  m, n, k = 200, 96, 54
  ag, bg, a, b = _operands(m, n, k, ta, tb, seed=5, pad_a=pad_a, pad_b=pad_b)
  c, status = eng.gemm_tf32(ag, bg, m, n, k, trans_a=ta, trans_b=tb)
  assert status == 0
  _close(c, _ref(a, b, ta, tb), EXACT_TOL, 'padded operands')


def test_dense_layer_epilogue():
  """y = x + relu(h) . W + b gated by a mask, then ReLU: every epilogue piece at once
  (ResNetDenseBlock.call, dense_resnet_value.py:63-69, and its reverse)."""
  from ibc_b200 import engine as eng
  m, n, k = 515, 200, 136
  g = torch.Generator().manual_seed(3)
  a, w = torch.randn(m, k, generator=g), torch.randn(k, n, generator=g) * 0.1
  bias, gate, res = torch.randn(n, generator=g), torch.randn(m, n, generator=g), \
      torch.randn(m, n, generator=g)
  c, status = eng.gemm_tf32(a.cuda(), w.cuda(), m, n, k, bias=bias.cuda(), gate_c=gate.cuda(),
                            res=res.cuda(), relu_a=True, relu_out=True)
  assert status == 0
  want = _ref(a, w, False, False, relu_a=True) + bias.double()
  want = torch.where(gate > 0, want, torch.zeros_like(want)) + res.double()
  _close(c, want.clamp_min(0), EXACT_TOL, 'fused epilogue')


@pytest.mark.parametrize('m,n,k', [(54, 32, 43200), (288, 64, 21600), (1152, 256, 1320),
                                   (258, 1024, 4113)])
def test_weight_gradient_form_accumulates_over_a_k_split(m, n, k):
  """dW += X^T . G (both operands contiguous along M / N, K = rows, split over CTAs)."""
  from ibc_b200 import engine as eng
  ag, bg, a, b = _operands(m, n, k, True, False, seed=k)
  g = torch.Generator().manual_seed(1)
  c0 = torch.randn(m, n, generator=g)
  c = c0.cuda().clone()
  _, status = eng.gemm_tf32(ag, bg, m, n, k, trans_a=True, out=c)
  assert status == 0
  _close(c, c0.double() + _ref(a, b, True, False), 2 * EXACT_TOL, 'split-K accumulation')


def test_pixel_engine_tf32_matches_its_fp32_path():
  """The PixelEBM handle computes the same encoder / value head in both precisions."""
  from ibc_b200 import _lib
  from ibc_b200 import engine as eng
  from oracle import pixel_ebm as opix
  spec = opix.PixelSpec(seq_len=2, height=40, width=56, channels=3, target_height=32,
                        target_width=48, act_dim=2, value_width=64, value_blocks=2)
  e = eng.PixelEngine(spec.seq_len, spec.height, spec.width, spec.channels, spec.target_height,
                      spec.target_width, spec.act_dim, spec.value_width, spec.value_blocks)
  assert e.precision == _lib.PRECISION_TF32
  flat = opix.init_params(spec, seed=1).cuda().contiguous()
  e.bind_params(flat)
  g = torch.Generator().manual_seed(0)
  img = torch.randint(0, 256, (5, 2, 40, 56, 3), generator=g, dtype=torch.uint8).cuda()
  enc_t = e.encode(img)
  assert e.kernel_status() == 0
  e.set_precision(_lib.PRECISION_FP32)
  enc_f = e.encode(img)
  _close(enc_t, enc_f, TF32_TOL, 'encoder tf32 vs fp32')


# ---- implicit 3x3 convolution (the A operand gathered from the NHWC tensor) --------------
CONV_SHAPES = [
    (2, 9, 13, 8, 32),        # vector path, ragged frame, BN = 32
    (3, 22, 30, 64, 128),     # conv4-like
    (1, 45, 60, 32, 64),      # conv2-like, odd height
    (2, 7, 5, 6, 32),         # 6 channels: element-wise gather (un-padded first layer)
    (1, 4, 6, 128, 256),      # frame smaller than a K-slice's row walk
]


def _conv_ref(x, k, bias=None, relu=False):
  import torch.nn.functional as F
  y = F.conv2d(round_tf32(x).double().permute(0, 3, 1, 2),
               round_tf32(k).double().permute(3, 2, 0, 1), None, padding=1).permute(0, 2, 3, 1)
  if bias is not None:
    y = y + bias.double()
  return y.clamp_min(0) if relu else y


@pytest.mark.parametrize('nb,h,w,c,co', CONV_SHAPES)
def test_implicit_convolution_forward_and_gradients(nb, h, w, c, co):
  """Keras Conv2D(3x3, 'same') (conv_maxpool.py:20-48): forward with fused bias + ReLU, data
  gradient (convolution with the flipped kernel) and weight gradient (split over rows,
  accumulated) against PyTorch conv2d / autograd in fp64 on the tf32-rounded operands."""
  from ibc_b200 import engine as eng
  g = torch.Generator().manual_seed(nb * 1000 + h * 10 + c)
  x = torch.randn(nb, h, w, c, generator=g)
  k = torch.randn(3, 3, c, co, generator=g) * 0.1
  bias = torch.randn(co, generator=g)
  dy = torch.randn(nb, h, w, co, generator=g)
  y, st = eng.conv3x3_tf32(x.cuda(), k.cuda(), bias=bias.cuda(), relu=True, mode=0)
  assert st == 0
  _close(y, _conv_ref(x, k, bias, relu=True), EXACT_TOL, 'forward')
  # gradients of sum(conv(x, k) * dy) with every MMA operand rounded
  xr = round_tf32(x).double().requires_grad_(True)
  kr = round_tf32(k).double().requires_grad_(True)
  import torch.nn.functional as F
  yy = F.conv2d(xr.permute(0, 3, 1, 2), kr.permute(3, 2, 0, 1), None, padding=1).permute(0, 2, 3, 1)
  gx, gk = torch.autograd.grad((yy * round_tf32(dy).double()).sum(), (xr, kr))
  dk0 = torch.randn(3, 3, c, co, generator=g)
  dk, st = eng.conv3x3_tf32(x.cuda(), k.cuda(), mode=1, dy=dy.cuda(), out=dk0.cuda().clone())
  assert st == 0
  _close(dk, dk0.double() + gk, 2 * EXACT_TOL, 'weight gradient')
  dx, st = eng.conv3x3_tf32(None, k.cuda(), mode=2, dy=dy.cuda())
  assert st == 0
  _close(dx, gx, EXACT_TOL, 'data gradient')
