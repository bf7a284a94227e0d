"""Differentiable wrappers: the reference's methods are @tf.function graphs of differentiable TensorFlow ops
(mdctransformer.py:61, :127; psychoacoustic.py:102, :122, and the comment on the gradient at :311), so a drop-in has
to work inside a training graph as well.  These torch.autograd.Functions route the forward through the library's
kernels and the backward through

  * the MDCT kernels themselves: with an orthogonal window (sine, vorbis) the analysis matrix is 1 / sqrt(4N) times a
    matrix with orthonormal rows whose transpose, times sqrt(4N), is the synthesis - so
        d loss / d x = inverse_transform(d loss / d Y)[:, N:-N] / 4N        (adjoint of transform)
        d loss / d Y = 4N transform(d loss / d x_hat)[:, 1:-1]              (adjoint of inverse_transform);
  * ac_pa_tonality_backward_f32 / ac_pa_threshold_backward_f32 (csrc/backward_kernels.cu) for the masking model.

MDCTransformer.transform / inverse_transform and PsychoacousticModel.tonality / global_masking_threshold switch to
these automatically when a torch input requires grad (float32 only).
"""

import torch

from . import _capi
from ._tensors import stream_ptr


def _orthogonal(mdct):
  if mdct.window_type.lower() not in ("sine", "vorbis"):
    raise NotImplementedError("the backward pass of the MDCT is built for the orthogonal windows ('sine', 'vorbis')")


class _Transform(torch.autograd.Function):
  @staticmethod
  def forward(ctx, x, mdct):
    ctx.mdct = mdct
    with torch.no_grad():
      return mdct.transform(x.detach())

  @staticmethod
  def backward(ctx, grad_y):
    mdct = ctx.mdct
    _orthogonal(mdct)
    n = mdct.filters_n
    with torch.no_grad():
      g = mdct.inverse_transform(grad_y.contiguous())
    return g[:, n:-n] / (4.0 * n), None


class _InverseTransform(torch.autograd.Function):
  @staticmethod
  def forward(ctx, y, mdct):
    ctx.mdct = mdct
    with torch.no_grad():
      return mdct.inverse_transform(y.detach())

  @staticmethod
  def backward(ctx, grad_x):
    mdct = ctx.mdct
    _orthogonal(mdct)
    with torch.no_grad():
      g = mdct.transform(grad_x.contiguous())
    return g[:, 1:-1] * (4.0 * mdct.filters_n), None


class _Tonality(torch.autograd.Function):
  @staticmethod
  def forward(ctx, a, pa):
    ctx.pa = pa
    a = a.detach().contiguous()
    ctx.save_for_backward(a)
    with torch.no_grad():
      return pa.tonality(a)

  @staticmethod
  def backward(ctx, grad_ton):
    (a,) = ctx.saved_tensors
    b, m, _, c = a.shape
    grad_a = torch.empty_like(a)
    g = grad_ton.contiguous().float()
    with torch.cuda.device(a.device):
      _capi.check(_capi.lib().ac_pa_tonality_backward_f32(ctx.pa._plan(a.device), a.data_ptr(), g.data_ptr(), grad_a.data_ptr(),
                                                          b, m, c, stream_ptr(a.device)))
    return grad_a, None


class _MaskingThreshold(torch.autograd.Function):
  @staticmethod
  def forward(ctx, a, ton, pa, drown):
    ctx.pa, ctx.drown = pa, float(drown)
    a, ton = a.detach().contiguous(), ton.detach().contiguous()
    ctx.save_for_backward(a, ton)
    with torch.no_grad():
      return pa.global_masking_threshold(a, ton, drown=drown)

  @staticmethod
  def backward(ctx, grad_thr):
    a, ton = ctx.saved_tensors
    b, m, _, c = a.shape
    grad_a, grad_ton = torch.empty_like(a), torch.empty_like(ton)
    g = grad_thr.contiguous().float()
    with torch.cuda.device(a.device):
      _capi.check(_capi.lib().ac_pa_threshold_backward_f32(ctx.pa._plan(a.device), a.data_ptr(), ton.data_ptr(), ctx.drown,
                                                           g.data_ptr(), grad_a.data_ptr(), grad_ton.data_ptr(), b, m, c,
                                                           stream_ptr(a.device)))
    return grad_a, grad_ton, None, None


def wants_grad(*tensors):
  return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def transform(mdct, x):
  return _Transform.apply(x, mdct)


def inverse_transform(mdct, y):
  return _InverseTransform.apply(y, mdct)


def tonality(pa, a):
  return _Tonality.apply(a, pa)


def global_masking_threshold(pa, a, ton, drown):
  return _MaskingThreshold.apply(a, ton, pa, drown)
