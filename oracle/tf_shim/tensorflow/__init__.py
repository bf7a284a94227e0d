"""NumPy stand-in for the sliver of the TensorFlow API that korneelvdbroek/audiocodec touches.

TEST INFRASTRUCTURE ONLY (lives under oracle/): TensorFlow is not installable in this image, so this
module lets the *unmodified* reference sources under /root/reference be imported and executed in the
build container (tests/golden/make_golden.py, tests/test_reference_under_shim.py).  Nothing in the
product path (audiocodec_b200/) may import it.

Every function below restates the documented semantics of the TF op of the same name with NumPy/SciPy:
  * tensors are plain numpy arrays / numpy scalars, dtypes are numpy dtypes (tf.float32 is np.float32);
  * python scalars are "weak" (NEP 50), exactly like TF's python-scalar promotion, so fp32 graphs stay fp32;
  * tf.nn.convolution is the NWC cross-correlation TF documents (no kernel flip);
  * tf.signal.dct(type=3, norm='ortho') is SciPy's DCT-III with the same normalisation TF documents.

Call sites it has to serve: mdctransformer.py:58-59,111-125,138-151,169-190,199-229,238-255,289-297,
306-309,327-347,362-368 and psychoacoustic.py:42-69,83-84,113-118,138-148,165-167,185-208,219-230,
240-255,281-299,312-313,330-339 (reference line numbers).
"""

import contextlib
import types

import numpy as _np
import scipy.fft as _sfft

# ---------------------------------------------------------------------------------------------- dtypes
float16 = _np.float16
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64


class _BFloat16Sentinel:
  """Only used for membership tests (psychoacoustic.py:42); no arithmetic in bf16 is emulated."""

  def __repr__(self):
    return "tf.bfloat16(shim sentinel)"


bfloat16 = _BFloat16Sentinel()
Tensor = _np.ndarray
DType = type


def _dt(dtype):
  return None if dtype is None else _np.dtype(dtype)


# -------------------------------------------------------------------------------------- construction
def constant(value, dtype=None, shape=None):
  out = _np.asarray(value, dtype=_dt(dtype))
  if shape is not None:
    out = _np.broadcast_to(out, shape).copy()
  return out if out.ndim else out[()]


def convert_to_tensor(value, dtype=None):
  return _np.asarray(value, dtype=_dt(dtype))


def cast(x, dtype):
  out = _np.asarray(x).astype(_np.dtype(dtype))
  return out if out.ndim else out[()]


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001 - mirrors tf.range
  if limit is None:
    start, limit = 0, start
  if dtype is None:
    is_float = any(isinstance(v, (float, _np.floating)) for v in (start, limit, delta))
    dtype = _np.float32 if is_float else _np.int32
  return _np.arange(start, limit, delta).astype(_np.dtype(dtype))


def ones(shape, dtype=float32):
  return _np.ones(shape, dtype=_np.dtype(dtype))


def zeros(shape, dtype=float32):
  return _np.zeros(tuple(_np.asarray(shape).tolist()) if not isinstance(shape, int) else shape,
                   dtype=_np.dtype(dtype))


def linspace(start, stop, num):
  start = _np.asarray(start)
  stop = _np.asarray(stop)
  dt = _np.result_type(start, stop)
  if num == 1:
    return _np.asarray([start], dtype=dt)
  # TF: start + delta * range(num), delta = (stop - start) / (num - 1), all in the operand dtype
  delta = ((stop - start) / dt.type(num - 1)).astype(dt)
  out = (start + delta * _np.arange(num).astype(dt)).astype(dt)
  out[-1] = stop
  return out


# ------------------------------------------------------------------------------------------- shaping
def shape(x):
  return _np.asarray(_np.shape(x), dtype=_np.int32)


def reshape(x, shape):  # noqa: A002
  shape = [int(s) for s in _np.asarray(shape).reshape(-1)] if not isinstance(shape, (list, tuple)) \
    else [int(s) for s in shape]
  return _np.reshape(x, shape)


def transpose(x, perm=None):
  return _np.transpose(x, axes=perm)


def expand_dims(x, axis):
  return _np.expand_dims(x, axis)


def reverse(x, axis):
  return _np.flip(x, axis=tuple(axis))


def concat(values, axis):
  return _np.concatenate(values, axis=axis)


def stack(values, axis=0):
  return _np.stack(values, axis=axis)


def pad(x, paddings, constant_values=0):
  return _np.pad(x, _np.asarray(paddings).tolist(), constant_values=constant_values)


def broadcast_to(x, shape):  # noqa: A002
  return _np.broadcast_to(x, shape)


# ---------------------------------------------------------------------------------------- elementwise
sin = _np.sin
sinh = _np.sinh
asinh = _np.arcsinh
exp = _np.exp
sqrt = _np.sqrt
abs = _np.abs  # noqa: A001
maximum = _np.maximum
minimum = _np.minimum
divide = _np.divide


def pow(x, y):  # noqa: A001
  return _np.power(x, y)


def clip_by_value(x, clip_value_min, clip_value_max):
  return _np.minimum(_np.maximum(x, clip_value_min), clip_value_max)


def map_fn(fn, elems):
  return _np.stack([_np.asarray(fn(e)) for e in elems], axis=0)


# ----------------------------------------------------------------------------------------- reductions
def reduce_mean(x, axis=None, keepdims=False):
  return _np.mean(x, axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype)


def reduce_sum(x, axis=None, keepdims=False):
  return _np.sum(x, axis=axis, keepdims=keepdims, dtype=_np.asarray(x).dtype)


def reduce_max(x, axis=None, keepdims=False):
  return _np.max(x, axis=axis, keepdims=keepdims)


def einsum(equation, *operands):
  return _np.einsum(equation, *operands)


# -------------------------------------------------------------------------------------------- helpers
def function(fn=None, **_kwargs):
  if fn is None:
    return lambda f: f
  return fn


@contextlib.contextmanager
def name_scope(_name):
  yield


# ---------------------------------------------------------------------------------------- sub-modules
def _diag(v):
  return _np.diag(_np.asarray(v))


def _inv(a):
  return _np.linalg.inv(a)


linalg = types.SimpleNamespace(diag=_diag, inv=_inv)
math = types.SimpleNamespace(log=_np.log, exp=_np.exp, pow=pow, sqrt=_np.sqrt)


def _convolution(input, filters, padding="VALID"):  # noqa: A002
  """1-D NWC cross-correlation: out[b, n, k] = sum_t sum_q input[b, n + t, q] * filters[t, q, k]."""
  assert padding == "VALID"
  taps = filters.shape[0]
  width = input.shape[1] - taps + 1
  out = _np.zeros((input.shape[0], width, filters.shape[2]), dtype=_np.result_type(input, filters))
  for t in _np.arange(taps):
    out += _np.matmul(input[:, t:t + width, :], filters[t])
  return out


nn = types.SimpleNamespace(convolution=_convolution)


def _dct(x, type=2, axis=-1, norm=None, n=None):  # noqa: A002
  x = _np.asarray(x)
  return _sfft.dct(x, type=type, n=n, axis=axis, norm=norm).astype(x.dtype)


signal = types.SimpleNamespace(dct=_dct)

_rng = _np.random.default_rng(20240607)


def _normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):  # noqa: A002
  g = _rng if seed is None else _np.random.default_rng(seed)
  return (mean + stddev * g.standard_normal(size=tuple(shape))).astype(_np.dtype(dtype))


def _uniform(shape, minval=0.0, maxval=1.0, dtype=float32, seed=None):  # noqa: A002
  g = _rng if seed is None else _np.random.default_rng(seed)
  return g.uniform(minval, maxval, size=tuple(shape)).astype(_np.dtype(dtype))


random = types.SimpleNamespace(normal=_normal, uniform=_uniform)
