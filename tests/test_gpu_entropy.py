"""The entropy-coded bitstream (ac_entropy_*_i32; SURVEY.md 8f row 4, no reference symbol): adaptive Golomb-Rice with
one independent byte range per frame.  Bit-exact against the CPU restatement (oracle/entropy_oracle.py), lossless on
the whole cfg2 tensor."""

import numpy as np
import pytest
import torch

import audiocodec_b200
from oracle import audiocodec_oracle as oracle
from oracle import entropy_oracle

pytestmark = pytest.mark.gpu


def _codec():
  return audiocodec_b200.AudioCodec(44100, filters_n=256)


@pytest.mark.parametrize("shape,scale", [((2, 3, 16, 1), 1.0), ((1, 5, 64, 2), 0.4), ((3, 2, 256, 2), 3.0),
                                         ((1, 4, 32, 3), 200.0), ((2, 2, 128, 1), 1e6), ((1, 1, 16, 1), 0.0)])
def test_stream_is_bit_exact_against_the_cpu_restatement(shape, scale):
  rng = np.random.default_rng(int(scale * 10) + shape[2])
  b, f, n, c = shape
  q = np.rint(rng.standard_normal(shape) * scale * rng.choice([0.0, 0.2, 1.0, 5.0], (b, f, 1, 1))).astype(np.int32)
  codec = _codec()
  stream, offsets = codec.pack(torch.from_numpy(q).cuda())
  ref_stream, ref_offsets = entropy_oracle.encode(q.reshape(b * f, n * c))
  assert np.array_equal(offsets.cpu().numpy(), ref_offsets)
  assert stream.numel() == ref_offsets[-1] + 4
  assert np.array_equal(stream.cpu().numpy()[:-4], ref_stream)
  back = codec.unpack(stream, offsets, shape)
  assert torch.equal(back.cpu(), torch.from_numpy(q))
  assert np.array_equal(entropy_oracle.decode(stream.cpu().numpy(), ref_offsets, b * f, n * c).reshape(shape), q)
  assert codec.stream_bytes(torch.from_numpy(q).cuda()) == ref_offsets[-1]


def test_extreme_values_and_long_runs():
  q = np.zeros((1, 4, 32, 1), dtype=np.int32)
  q[0, 0, 3] = 2 ** 31 - 1
  q[0, 1, 0] = -2 ** 31
  q[0, 1, 17:] = 1
  q[0, 2, :] = -70000
  q[0, 2, 5] = 1 << 20                       # an outlier next to small values: a unary run of many zero words
  codec = _codec()
  stream, offsets = codec.pack(torch.from_numpy(q).cuda())
  ref_stream, ref_offsets = entropy_oracle.encode(q.reshape(4, 32))
  assert np.array_equal(offsets.cpu().numpy(), ref_offsets) and np.array_equal(stream.cpu().numpy()[:-4], ref_stream)
  assert torch.equal(codec.unpack(stream, offsets, q.shape).cpu(), torch.from_numpy(q))
  assert int(offsets[4] - offsets[3]) == 4   # an all-zero frame: two 5-bit headers in one word
  with pytest.raises(ValueError):
    codec.pack(torch.zeros(1, 1, 8, 1, dtype=torch.int32, device="cuda"))      # row_len not a multiple of 16


def test_full_size_cfg2_is_lossless():
  """All of cfg2 through encode -> pack -> unpack -> decode: the integers come back bit for bit, the row sizes equal
  the CPU restatement's, and the stream takes 2 - 3 bits per coefficient."""
  sr, n = 44100, 256
  s = (sr * 10 // n) * n
  codec = _codec()
  x = torch.from_numpy(oracle.synthetic_audio(64, s, 2, sr)).cuda()
  q, step = codec.encode(x)
  stream, offsets = codec.pack(q)
  assert torch.equal(codec.unpack(stream, offsets, q.shape), q)
  sizes = (offsets[1:] - offsets[:-1]).cpu().numpy()
  rows = q.shape[0] * q.shape[1]
  probe = np.arange(0, rows, 97)
  assert np.array_equal(sizes[probe], entropy_oracle.row_sizes(q.reshape(rows, -1)[probe].cpu().numpy()))
  est = codec.psychoacoustic.bit_estimate(q)
  total = int(offsets[-1].item())
  assert stream.numel() == total + 4
  bits_per_coef = 8.0 * total / q.numel()
  # a Rice code spends at least one bit per coefficient: cfg2 (half of the integers zero, mean |q| 0.84) measures 2.4
  # bits against the 1.04 of the estimate sum log2(2|q| + 1) / n the rate loop steers by, and 32 of the int32 tensor
  assert 1.0 < bits_per_coef < 3.0, bits_per_coef
  assert est["bits"] / q.numel() < bits_per_coef
  st = codec.stats(q)
  assert st[3].item() == total and st[0].item() == q.numel()
