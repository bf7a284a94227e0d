"""ctypes binding of include/audiocodec_b200.h (the C ABI of the CUDA library).

The library is loaded from audiocodec_b200/lib/ (built in-tree by audiocodec_b200.build).  There is no
fallback of any kind: a missing library is an ImportError-grade failure, a missing GPU surfaces as the
library's AC_ERR_CUDA on the first plan creation.
"""

import ctypes
import os

from . import build as _build

AC_OK, AC_ERR_INVALID, AC_ERR_CUDA, AC_ERR_UNSUPPORTED, AC_ERR_ALLOC = 0, -1, -2, -3, -4
DTYPE_F32, DTYPE_BF16 = 0, 1
WINDOW_ONES, WINDOW_SINE, WINDOW_VORBIS = 0, 1, 2

_c_int64 = ctypes.c_int64
_c_void_p = ctypes.c_void_p
_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_double_p = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); mirrors the header one to one (tests/test_host_side.py::test_library_exports_every_declared_symbol cross-checks the names)
SIGNATURES = {
  "ac_last_error": (ctypes.c_char_p, []),
  "ac_abi_version": (ctypes.c_int, []),
  "ac_kernel_launch_count": (_c_int64, []),
  "ac_mdct_tables_host": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p]),
  "ac_pa_tables_host": (ctypes.c_int, [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       _c_float_p, _c_float_p, _c_float_p, _c_float_p, _c_double_p]),
  "ac_pa_mma_jobs_host": (ctypes.c_int, [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double, _c_void_p,
                                         _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
  "ac_mdct_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_c_void_p)]),
  "ac_mdct_plan_create_ex": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_c_void_p)]),
  "ac_mdct_plan_destroy": (ctypes.c_int, [_c_void_p]),
  "ac_mdct_forward_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_mdct_inverse_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_mdct_inverse_dequant_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64,
                                                 ctypes.c_int, _c_void_p]),
  "ac_mdct_inverse_dequant_compact_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, ctypes.c_float,
                                                         _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_plan_create": (ctypes.c_int, [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                       ctypes.POINTER(_c_void_p)]),
  "ac_pa_plan_create_ex": (ctypes.c_int, [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                                          ctypes.POINTER(_c_void_p)]),
  "ac_pa_plan_destroy": (ctypes.c_int, [_c_void_p]),
  "ac_pa_tonality_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_threshold_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_int64,
                                         _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_encode_f32": (ctypes.c_int, [_c_void_p, _c_void_p, ctypes.c_float, ctypes.c_float, _c_void_p, _c_void_p,
                                      _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_encode_compact_f32": (ctypes.c_int, [_c_void_p, _c_void_p, ctypes.c_float, ctypes.c_float, _c_void_p, _c_void_p,
                                              _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_expand_threshold_f32": (ctypes.c_int, [_c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_int64, _c_int64,
                                                ctypes.c_int, _c_void_p]),
  "ac_codec_encode_workspace_bytes": (_c_int64, [_c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int]),
  "ac_codec_encode_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_float, ctypes.c_float, _c_void_p,
                                         _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p, _c_void_p]),
  "ac_pa_add_noise_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, ctypes.c_uint64, _c_void_p]),
  "ac_quantize_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_void_p]),
  "ac_dequantize_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_void_p]),
  "ac_codec_stats_i32": (ctypes.c_int, [_c_void_p, _c_int64, _c_void_p, _c_void_p]),
  "ac_mdct_forward_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_mdct_inverse_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_tonality_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_threshold_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_double, _c_void_p, _c_int64,
                                         _c_int64, ctypes.c_int, _c_void_p]),
  "ac_quantize_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_void_p]),
  "ac_dequantize_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_void_p]),
  "ac_pa_tonality_backward_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int,
                                                 _c_void_p]),
  "ac_pa_threshold_backward_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_void_p,
                                                  _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_entropy_plan_i32": (ctypes.c_int, [_c_void_p, _c_int64, _c_int64, _c_void_p, _c_void_p]),
  "ac_entropy_encode_i32": (ctypes.c_int, [_c_void_p, _c_int64, _c_int64, _c_void_p, _c_void_p, _c_void_p]),
  "ac_entropy_decode_i32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_int64, _c_int64, _c_void_p, _c_void_p]),
  "ac_bf16_workspace_bytes": (_c_int64, [_c_int64, _c_int64]),
  "ac_mdct_forward_bf16": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p, _c_void_p]),
  "ac_mdct_inverse_bf16": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p, _c_void_p]),
  "ac_pa_tonality_bf16": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int, _c_void_p, _c_void_p]),
  "ac_pa_threshold_bf16": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_int64, _c_int64,
                                          ctypes.c_int, _c_void_p, _c_void_p]),
  "ac_pa_amplitude_to_db_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_pa_amplitude_to_db_f64": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, ctypes.c_int, _c_void_p]),
  "ac_codec_pipeline_create": (ctypes.c_int, [_c_void_p, _c_void_p, _c_int64, _c_int64, ctypes.c_int,
                                              ctypes.POINTER(_c_void_p)]),
  "ac_codec_pipeline_destroy": (ctypes.c_int, [_c_void_p]),
  "ac_codec_roundtrip_host_f32": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int64, ctypes.c_float,
                                                 ctypes.c_float, _c_double_p, _c_void_p]),
  "ac_mdct_forward_dl": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p]),
  "ac_mdct_inverse_dl": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p]),
  "ac_pa_tonality_dl": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p]),
  "ac_pa_threshold_dl": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p, ctypes.c_float, _c_void_p, _c_void_p]),
}

_lib = None


def library_path():
  # AUDIOCODEC_B200_LIB: an alternative build of the same library (A/B runs of kernel variants, tools/)
  return os.environ.get("AUDIOCODEC_B200_LIB") or _build.lib_path()


def lib():
  """The loaded shared library (loads on first use; raises if it has not been built)."""
  global _lib
  if _lib is None:
    path = library_path()
    if not os.path.exists(path):
      raise ImportError(
        f"{path} is missing: build the CUDA library first (python -m audiocodec_b200.build). "
        "audiocodec_b200 has no CPU or pure-PyTorch fallback.")
    handle = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(handle, name)
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = handle
  return _lib


def check(rc):
  """Maps an ac_status to the exception the reference raises for the same mistake."""
  if rc == AC_OK:
    return
  msg = lib().ac_last_error().decode("utf-8", "replace")
  if rc == AC_ERR_INVALID:
    raise ValueError(msg)
  if rc == AC_ERR_UNSUPPORTED:
    raise NotImplementedError(msg)
  if rc == AC_ERR_ALLOC:
    raise MemoryError(msg)
  raise RuntimeError(msg)


# ---- DLPack capsules -------------------------------------------------------------------------------------
_PyCapsule_GetPointer = ctypes.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = ctypes.c_void_p
_PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_PyCapsule_IsValid = ctypes.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = ctypes.c_int
_PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]


def dl_pointer(capsule):
  """DLManagedTensor* held by a (not yet consumed) "dltensor" capsule; the capsule keeps ownership."""
  if not _PyCapsule_IsValid(capsule, b"dltensor"):
    raise ValueError("expected an unconsumed DLPack capsule named 'dltensor'")
  return _PyCapsule_GetPointer(capsule, b"dltensor")
