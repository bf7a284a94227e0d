"""Tensor plumbing for the host-side classes: device memory, streams and DLPack hand-over via PyTorch.

PyTorch is used for allocation and stream bookkeeping only; every computation goes through the C ABI.
Foreign tensors (anything exposing __dlpack__, e.g. TensorFlow eager tensors) are viewed zero-copy and
results are handed back in the caller's framework.  The kernels need C-contiguous, 16-byte aligned tensors: a
C-contiguous aligned input is passed as is (no copy); a strided view or a mis-aligned slice is compacted into a
fresh device buffer first - one extra device-to-device pass, made visible by `adopt.copies` (tests assert it stays 0
on the hot path).  The raw C ABI and the *_dl entry points never copy: they refuse such tensors with AC_ERR_INVALID.
"""

import numpy as np
import torch


def normalise_compute_dtype(dtype, who):
  """Accepts tf / torch / numpy dtypes or strings: float32 (the tuned path), float64 and bfloat16 (functional paths)."""
  name = getattr(dtype, "name", None) or str(dtype)
  name = name.replace("torch.", "").replace("<dtype: '", "").replace("'>", "")
  if name in ("float32", "float", "f32") or dtype is np.float32:
    return "float32"
  if name in ("float64", "double", "f64") or dtype is np.float64:
    return "float64"
  if name in ("bfloat16", "bf16"):
    return "bfloat16"
  raise TypeError(f"compute_dtype of {who} should be float64, float32 or bfloat16 (got {dtype!r})")


def normalise_precompute_dtype(dtype):
  name = getattr(dtype, "name", None) or str(dtype)
  name = name.replace("torch.", "").replace("<dtype: '", "").replace("'>", "")
  if name in ("float64", "double") or dtype is np.float64:
    return "float64"
  if name in ("float32", "float") or dtype is np.float32:
    return "float32"
  raise TypeError(f"precompute_dtype must be float64 or float32 (got {dtype!r})")


def adopt(x, name, dtype=torch.float32):
  """Returns (torch view of x, function mapping a torch result back to x's framework)."""
  back = lambda t: t  # noqa: E731
  if not isinstance(x, torch.Tensor):
    if not hasattr(x, "__dlpack__"):
      raise TypeError(f"{name}: expected a device tensor (torch.Tensor or any object with __dlpack__), got {type(x)!r}")
    module = type(x).__module__
    x = torch.from_dlpack(x)
    if module.startswith("tensorflow"):
      def back(t):
        import tensorflow as tf  # lazy: only when the caller is TensorFlow
        return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t))
  if x.dtype != dtype:
    raise TypeError(f"{name}: dtype {x.dtype} does not match the compute dtype {dtype} (no implicit casting)")
  if not x.is_cuda:
    raise RuntimeError(f"{name}: tensor is on {x.device}; audiocodec_b200 only runs on CUDA devices (no CPU path)")
  if not x.is_contiguous() or x.data_ptr() % 16 != 0:
    adopt.copies += 1              # the one place a tensor is copied: strided or mis-aligned input (see module docstring)
    x = x.contiguous() if not x.is_contiguous() else x.clone()
    if x.data_ptr() % 16 != 0:
      x = x.clone()
  return x, back


adopt.copies = 0


def torch_dtype(name):
  return {"float64": torch.float64, "bfloat16": torch.bfloat16}.get(name, torch.float32)


def bf16_workspace(lib, in_elems, out_elems, device):
  """float32 scratch of the bfloat16 entry points (the kernels run in float32 on copies of the bfloat16 tensors)."""
  need = lib.ac_bf16_workspace_bytes(int(in_elems), int(out_elems))
  return torch.empty(max(need // 4, 4), dtype=torch.float32, device=device)


def stream_ptr(device):
  return torch.cuda.current_stream(device).cuda_stream
