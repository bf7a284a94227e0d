"""The single-pass encoder (ac_codec_encode_f32): x -> (q, step | bark thresholds) in one kernel for
stereo signals with filters_n = 256, the amplitudes never in global memory; other shapes through a workspace.

It must be BIT-IDENTICAL to the two-kernel chain transform -> encode (same operations in the same order; the oracle
parity of that chain is tests/test_gpu_mdct.py / test_gpu_psycho.py / test_gpu_full_size.py - the last one runs
AudioCodec.encode, i.e. this kernel, over all of cfg2 and slices of cfg5)."""

import numpy as np
import pytest
import torch

import audiocodec_b200
from audiocodec_b200 import _capi
from oracle import audiocodec_oracle as oracle

pytestmark = pytest.mark.gpu


def cuda(a):
  return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("b,blocks", [(1, 1), (2, 3), (1, 30), (1, 31), (2, 32), (1, 33), (3, 40), (2, 100), (1, 0), (5, 64)])
def test_fused_equals_two_kernel_chain(b, blocks):
  """Ragged tiles on both sides of the 32-frame tile, one and several batch rows, an empty signal."""
  n, sr = 256, 44100
  x = cuda(oracle.synthetic_audio(b, blocks * n, 2, sr))
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  lib = _capi.lib()
  assert lib.ac_codec_encode_workspace_bytes(codec.mdct._plan(x.device), codec.psychoacoustic._plan(x.device), b, blocks * n, 2) == 0
  y = codec.mdct.transform(x)
  q_ref, step_ref = codec.psychoacoustic.encode(y)
  launches0 = lib.ac_kernel_launch_count()
  q, step = codec.encode(x)
  assert lib.ac_kernel_launch_count() - launches0 == (1 if b * (blocks + 1) > 0 else 0)     # ONE kernel
  assert q.dtype == torch.int32 and tuple(q.shape) == (b, blocks + 1, n, 2)
  assert torch.equal(q, q_ref)
  assert torch.equal(step, step_ref)
  # q alone, and the fixed-bitrate scalar / drown of the reference's signature
  assert torch.equal(codec.encode(x, return_threshold=False), q_ref)
  q2, step2 = codec.encode(x, drown=0.35, thr_scale=1.7)
  q2_ref, step2_ref = codec.psychoacoustic.encode(y, drown=0.35, thr_scale=1.7)
  assert torch.equal(q2, q2_ref) and torch.equal(step2, step2_ref)


def test_fused_compact_side_information():
  n, sr = 256, 44100
  x = cuda(oracle.synthetic_audio(3, 45 * n, 2, sr))
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  y = codec.mdct.transform(x)
  q_ref, g_ref = codec.psychoacoustic.encode_compact(y)
  q, g = codec.encode(x, compact=True)
  assert tuple(g.shape) == (3, 46, 64, 2)
  assert torch.equal(q, q_ref) and torch.equal(g, g_ref)
  q_s, step = codec.encode(x)
  assert torch.equal(q_s, q)
  assert torch.equal(codec.decode_compact(q, g), codec.decode(q, step))      # the decoder rebuilds the same steps


def test_fused_repeated_runs_and_other_streams():
  """The T region is block rows, FFT scratch and transposed amplitudes in turn, and the ticket counter hands out the
  tiles: twenty runs, some on a side stream, must agree bit for bit."""
  n, sr = 256, 44100
  x = cuda(oracle.synthetic_audio(8, 301 * n, 2, sr))
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  q0, s0 = codec.encode(x)
  side = torch.cuda.Stream()
  junk = torch.empty(1 << 22, device="cuda")
  for i in range(20):
    if i % 3 == 0:
      junk.normal_()
    if i % 4 == 1:
      side.wait_stream(torch.cuda.current_stream())
      with torch.cuda.stream(side):
        q, s = codec.encode(x)
      torch.cuda.current_stream().wait_stream(side)
    else:
      q, s = codec.encode(x)
    assert torch.equal(q, q0) and torch.equal(s, s0)


@pytest.mark.parametrize("sr,n,c,b,blocks", [(44100, 256, 1, 2, 37), (48000, 1024, 2, 1, 9), (44100, 256, 3, 2, 5),
                                            (44100, 64, 2, 2, 11), (44100, 100, 1, 1, 4)])
def test_other_shapes_run_through_the_workspace(sr, n, c, b, blocks):
  x = cuda(oracle.synthetic_audio(b, blocks * n, c, sr))
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  need = _capi.lib().ac_codec_encode_workspace_bytes(codec.mdct._plan(x.device), codec.psychoacoustic._plan(x.device), b, blocks * n, c)
  assert need == 4 * b * (blocks + 1) * n * c
  q, step = codec.encode(x)
  q_ref, step_ref = codec.psychoacoustic.encode(codec.mdct.transform(x))
  assert torch.equal(q, q_ref) and torch.equal(step, step_ref)
  with pytest.raises(ValueError):           # the C entry point refuses these shapes without a workspace
    _capi.check(_capi.lib().ac_codec_encode_f32(codec.mdct._plan(x.device), codec.psychoacoustic._plan(x.device),
                                                x.data_ptr(), 0.0, 1.0, step.data_ptr(), None, q.data_ptr(), b, blocks * n, c,
                                                None, torch.cuda.current_stream().cuda_stream))


def test_full_size_cfg2_fused_equals_chain():
  """All of cfg2 (64 clips, 112.9 M coefficients): the fused encoder and the two-kernel chain agree bit for bit."""
  sr, n = 44100, 256
  s = (sr * 10 // n) * n
  x = torch.rand(64, s, 2, device="cuda") - 0.5
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  q, step = codec.encode(x)
  q_ref, step_ref = codec.psychoacoustic.encode(codec.mdct.transform(x))
  assert torch.equal(q, q_ref) and torch.equal(step, step_ref)
