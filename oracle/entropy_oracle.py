"""CPU restatement of the entropy-coded bitstream (audiocodec_b200/csrc/entropy_kernels.cu).  TEST INFRASTRUCTURE ONLY.

The reference has no quantiser and no bitstream (SURVEY.md finding 2): the format is build-defined, this module is
its second, independent implementation (plain Python / NumPy) against which the CUDA coder is checked bit for bit.

Format: q is `rows` rows of `row_len` integers, groups of 16.  Per group a 5-bit Rice parameter k (31 = all zero, no
payload); each value zigzag-mapped u = (q << 1) ^ (q >> 31) and written as (u >> k) zero bits, a one bit, the k low bits
(least significant first).  k = argmin over {k0 - 1, k0, k0 + 1} (clipped to [0, 30], lowest on ties) of the payload
bits, k0 = floor(log2(sum(u) // 16 + 1)).  Bits fill little-endian 32-bit words from bit 0; a row is padded to whole
words, so rows start on 4-byte boundaries: offsets[r] (bytes), offsets[rows] = size of the stream.
"""

import numpy as np

GROUP = 16


def zigzag(q):
  q = np.asarray(q, dtype=np.int64)
  return ((q << 1) ^ (q >> 63)).astype(np.uint64) & np.uint64(0xffffffff)


def unzigzag(u):
  u = np.asarray(u, dtype=np.int64)
  return ((u >> 1) ^ -(u & 1)).astype(np.int32)


def _choose_k(u):
  """u [..., 16] uint64 -> (k [...], payload bits [...])."""
  total = u.sum(axis=-1)
  mean1 = total // GROUP + 1
  k0 = np.floor(np.log2(mean1.astype(np.float64))).astype(np.int64)
  # exact floor(log2) for integers (float rounding can be off by one at powers of two)
  k0 = np.where((np.int64(1) << np.minimum(k0 + 1, 62)) <= mean1.astype(np.int64), k0 + 1, k0)
  k0 = np.where((np.int64(1) << np.maximum(k0, 0)) > mean1.astype(np.int64), k0 - 1, k0)
  best_k = np.full(total.shape, -1, dtype=np.int64)
  best_bits = np.zeros(total.shape, dtype=np.int64)
  for d in (-1, 0, 1):
    k = k0 + d
    valid = (k >= max(0, 0)) & (k <= 30) & (k >= np.maximum(k0 - 1, 0)) & (k <= np.minimum(k0 + 1, 30))
    kk = np.clip(k, 0, 30)
    bits = ((u.astype(np.int64) >> kk[..., None]) + 1 + kk[..., None]).sum(axis=-1)
    take = valid & ((best_k < 0) | (bits < best_bits))
    best_k = np.where(take, kk, best_k)
    best_bits = np.where(take, bits, best_bits)
  zero = total == 0
  return np.where(zero, 31, best_k), np.where(zero, 0, best_bits)


def row_sizes(q2d):
  """Bytes per row (a multiple of four), vectorised: usable on whole tensors."""
  q2d = np.asarray(q2d)
  rows, row_len = q2d.shape
  assert row_len % GROUP == 0
  u = zigzag(q2d).reshape(rows, row_len // GROUP, GROUP)
  _, bits = _choose_k(u)
  total = (bits + 5).sum(axis=1)
  return ((total + 31) // 32) * 4


def encode(q2d):
  """-> (stream uint8 [bytes], offsets int64 [rows + 1]).  Python loops: small inputs only."""
  q2d = np.asarray(q2d)
  rows, row_len = q2d.shape
  sizes = row_sizes(q2d)
  offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
  stream = np.zeros(int(offsets[-1]), dtype=np.uint8)
  u_all = zigzag(q2d).reshape(rows, row_len // GROUP, GROUP)
  k_all, _ = _choose_k(u_all)
  for r in range(rows):
    acc, n = 0, 0                                   # a Python integer as the row's bit string, bit 0 first
    for g in range(row_len // GROUP):
      k = int(k_all[r, g])
      acc |= k << n
      n += 5
      if k == 31:
        continue
      for v in u_all[r, g]:
        v = int(v)
        n += v >> k                                 # zeros
        acc |= 1 << n
        n += 1
        if k:
          acc |= (v & ((1 << k) - 1)) << n
          n += k
    nbytes = int(sizes[r])
    assert n <= 8 * nbytes
    stream[offsets[r]:offsets[r] + nbytes] = np.frombuffer(acc.to_bytes(nbytes, "little"), dtype=np.uint8)
  return stream, offsets


def decode(stream, offsets, rows, row_len):
  out = np.zeros((rows, row_len), dtype=np.int32)
  for r in range(rows):
    acc = int.from_bytes(bytes(stream[offsets[r]:offsets[r + 1]]), "little")
    pos = 0
    for g in range(row_len // GROUP):
      k = (acc >> pos) & 31
      pos += 5
      if k == 31:
        continue
      for i in range(GROUP):
        z = 0
        while not (acc >> pos) & 1:
          pos += 1
          z += 1
        pos += 1
        u = z << k
        if k:
          u |= (acc >> pos) & ((1 << k) - 1)
          pos += k
        out[r, g * GROUP + i] = (u >> 1) ^ -(u & 1)
  return out
