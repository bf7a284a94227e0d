"""Backward pass: the drop-in classes as differentiable layers, as the reference's @tf.function
methods are (mdctransformer.py:61, :127; psychoacoustic.py:102, :122, gradient hint at :311).

  * MDCT: both operators are linear, so the adjoint (dot-product) identity <T x, g> = <x, T^T g> is the whole test;
  * masking model: vector-Jacobian products of the CUDA kernels against torch.autograd on a float64 restatement of the
    formulas (psychoacoustic.py:113-118, :139-146, :185-208, :312-313, :330-331) built from the model's own tables.
"""

import math

import numpy as np
import pytest
import torch

import audiocodec_b200
from oracle import audiocodec_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,window,b,blocks,c", [(256, 'vorbis', 2, 37, 2), (1024, 'sine', 1, 5, 2), (64, 'vorbis', 3, 9, 1),
                                                 (100, 'sine', 1, 4, 3)])
def test_mdct_adjoints(n, window, b, blocks, c):
  g0 = torch.Generator(device="cuda").manual_seed(n)
  mdct = audiocodec_b200.MDCTransformer(n, window_type=window)
  x = (torch.rand(b, blocks * n, c, device="cuda", generator=g0) - 0.5).requires_grad_()
  y = mdct.transform(x)
  assert y.requires_grad and tuple(y.shape) == (b, blocks + 1, n, c)
  g = torch.randn(y.shape, device="cuda", generator=g0)
  (y * g).sum().backward()
  lhs = (y.detach().double() * g.double()).sum().item()
  rhs = (x.detach().double() * x.grad.double()).sum().item()
  assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1e-3), (lhs, rhs)
  # inverse_transform
  a = torch.randn(b, blocks + 1, n, c, device="cuda", generator=g0).requires_grad_()
  xh = mdct.inverse_transform(a)
  h = torch.randn(xh.shape, device="cuda", generator=g0)
  (xh * h).sum().backward()
  lhs = (xh.detach().double() * h.double()).sum().item()
  rhs = (a.detach().double() * a.grad.double()).sum().item()
  assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1e-3), (lhs, rhs)
  # a round trip is the identity away from the edges: so is its Jacobian
  x2 = (torch.rand(b, blocks * n, c, device="cuda", generator=g0) - 0.5).requires_grad_()
  w = torch.randn(b, blocks * n, c, device="cuda", generator=g0)
  (mdct.inverse_transform(mdct.transform(x2))[:, n:-n] * w).sum().backward()
  assert (x2.grad - w).abs().max().item() < 2e-5 * w.abs().max().item()
  with pytest.raises(NotImplementedError):
    z = torch.zeros(1, 2 * n, 1, device="cuda", requires_grad=True)
    audiocodec_b200.MDCTransformer(n, window_type='ones').transform(z).sum().backward()


def _ref_tonality(a, eps):
  i = a ** 2
  log_gm = torch.log(torch.clamp_min(i, eps)).mean(dim=2, keepdim=True)
  am = i.mean(dim=2, keepdim=True) + eps
  sfm = 10.0 * (log_gm - torch.log(am)) / math.log(10.0)
  return torch.clamp_max(sfm / -60.0, 1.0)


def _ref_threshold(a, ton, pa, drown):
  eps = 1e-14
  dev = a.device
  w = pa.W.double().to(dev)                   # [N, nb]
  w_inv = pa.W_inv.double().to(dev)           # [nb, N]
  s = pa.spreading_matrix.double().to(dev)    # [nb, nb]
  quiet = pa.quiet_threshold_intensity.double().to(dev)
  nb = pa.bark_bands_n
  lin = torch.linspace(0.0, float(np.float32(pa.max_bark)), nb, dtype=torch.float32).double().to(dev)
  offset = (1.0 - drown) * (torch.einsum('nbic,j->nbjc', ton, lin) + 9.0 * ton + 5.5)
  gain = torch.pow(torch.tensor(10.0, dtype=torch.float64, device=dev), -pa.alpha * offset / 10.0)
  bark = torch.einsum('nbic,ij->nbjc', a ** 2, w)
  p = torch.clamp_min(bark, eps) ** pa.alpha
  m = torch.einsum('nbic,ij->nbjc', p, s) * gain
  mk = torch.clamp_min(m, eps) ** (1.0 / pa.alpha)
  g = torch.maximum(mk, quiet)
  v = torch.einsum('nbic,ij->nbjc', g, w_inv)
  return torch.sqrt(torch.clamp_min(v, eps))


@pytest.mark.parametrize("sr,n,c,drown", [(44100, 256, 2, 0.0), (48000, 1024, 1, 0.35), (16000, 64, 3, 0.0)])
def test_masking_model_gradients(sr, n, c, drown):
  x = oracle.synthetic_audio(2, 12 * n, c, sr)
  y = oracle.MDCTransformer(n).transform(x)
  y[0, 0] = 0.0                                             # an all-zero frame: every clamp active
  pa = audiocodec_b200.PsychoacousticModel(sr, n)
  a = torch.from_numpy(y).cuda().requires_grad_()
  ton = pa.tonality(a)
  assert ton.requires_grad
  g_ton = torch.randn(ton.shape, device="cuda")
  ton.backward(g_ton)
  a64 = torch.from_numpy(y).cuda().double().requires_grad_()
  ton64 = _ref_tonality(a64, 1e-14)
  ton64.backward(g_ton.double())
  scale = a64.grad.abs().max().item()
  assert (a.grad.double() - a64.grad).abs().max().item() <= 2e-4 * scale

  # threshold: gradients with respect to the amplitudes AND the tonality input
  a = torch.from_numpy(y).cuda().requires_grad_()
  t_in = ton.detach().clone().requires_grad_()
  thr = pa.global_masking_threshold(a, t_in, drown=drown)
  g_thr = torch.randn(thr.shape, device="cuda")
  thr.backward(g_thr)
  a64 = torch.from_numpy(y).cuda().double().requires_grad_()
  t64 = ton.detach().double().requires_grad_()
  thr64 = _ref_threshold(a64, t64, pa, drown)
  assert (thr.detach().double() - thr64.detach()).abs().max().item() <= 2e-4 * thr64.abs().max().item()
  thr64.backward(g_thr.double())
  # compare per frame relative to the frame's largest gradient (fp32 kernels against a float64 graph)
  da, da64 = a.grad.double(), a64.grad
  frame_scale = da64.abs().amax(dim=2, keepdim=True).clamp_min(1e-12)
  assert ((da - da64).abs() / frame_scale).max().item() <= 2e-3
  dt, dt64 = t_in.grad.double(), t64.grad
  assert ((dt - dt64).abs() / dt64.abs().clamp_min(1e-6 * dt64.abs().max())).max().item() <= 2e-3

  # the whole chain as a layer: loss = sum(thr(transform(x))) differentiates back to the signal
  mdct = audiocodec_b200.MDCTransformer(n)
  xs = torch.from_numpy(x).cuda().requires_grad_()
  yy = mdct.transform(xs)
  loss = pa.global_masking_threshold(yy, pa.tonality(yy)).sum()
  loss.backward()
  assert xs.grad is not None and torch.isfinite(xs.grad).all() and xs.grad.abs().max().item() > 0
