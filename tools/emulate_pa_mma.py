"""NumPy emulation of the index maps of pa_mma_tile_kernel's spreading phase (psycho_mma_kernels.cu).

Checks, without a GPU, that the XOR-swizzled P layout, the mma.m16n8k8 fragment addressing, the rotating Toeplitz
B fragments and the accumulator -> G scatter compute G[j][item] = sum_i P[item][i] * spread_fn[64 - i + j], and that
the shared-memory accesses involved are bank-conflict free.
"""
import numpy as np

TI, GS, NB = 64, 68, 64        # items per tile, row stride of G, bark bands (psycho_mma_kernels.cu)


def mma_m16n8k8(acc, a, b):
  """acc[lane][4] += A B with the PTX fragment layout; a[lane][4], b[lane][2]."""
  A = np.zeros((16, 8))
  B = np.zeros((8, 8))
  for lane in range(32):
    g, t = lane >> 2, lane & 3
    A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
    B[t, g], B[t + 4, g] = b[lane]
  D = A @ B
  for lane in range(32):
    g, t = lane >> 2, lane & 3
    acc[lane] += (D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1])


def banks_ok(addresses):
  """one wavefront: all 32 word addresses in distinct banks, or equal addresses (broadcast)"""
  by_bank = {}
  for a in addresses:
    by_bank.setdefault(a % 32, set()).add(a)
  return all(len(v) == 1 for v in by_bank.values())


def main():
  rng = np.random.default_rng(0)
  p_true = rng.random((TI, NB)) + 0.1          # [item][band]
  sf = rng.random(128) + 0.1
  # A2 store: lane l owns the item pair (2l, 2l + 1): one 64-bit store at P[band * TI + ((2l) ^ ((band & 3) << 3))]
  P = np.zeros(NB * TI)
  for band in range(NB):
    addrs = [band * TI + ((2 * lane) ^ ((band & 3) << 3)) for lane in range(32)]
    for half in (addrs[:16], addrs[16:]):          # a 64-bit access is served per half-warp
      assert banks_ok(half) and banks_ok([a + 1 for a in half])
    for lane in range(32):
      P[addrs[lane]] = p_true[2 * lane, band]
      P[addrs[lane] + 1] = p_true[2 * lane + 1, band]
  G = np.full(NB * GS, np.nan)
  for warp in range(8):
    m0, nq = (warp & 3) * 16, warp >> 2
    acc = np.zeros((4, 32, 4))
    b = [None] * 4
    lanes = range(32)
    gg = [l >> 2 for l in lanes]
    tt = [l & 3 for l in lanes]
    lb = [64 + gg[l] - tt[l] + 32 * nq for l in lanes]
    for nt in range(4):
      idx = [lb[l] + 8 * nt for l in lanes]
      assert min(idx) - 4 >= 0 and max(idx) < 128
      assert banks_ok(idx) and banks_ok([i - 4 for i in idx])
      b[nt] = np.array([(sf[i], sf[i - 4]) for i in idx])
    for ks in range(8):
      if ks > 0:
        for nt in (3, 2, 1):
          b[nt] = b[nt - 1]
        idx = [lb[l] - 8 * ks for l in lanes]
        assert min(idx) - 4 >= 0 and max(idx) < 128
        b[0] = np.array([(sf[i], sf[i - 4]) for i in idx])
      a = np.zeros((32, 4))
      for e, (drow, dcol) in enumerate(((0, 0), (0, 8), (4, 0), (4, 8))):
        addrs = []
        for l in lanes:
          col = (m0 + gg[l] + dcol) ^ (tt[l] << 3)
          addrs.append((8 * ks + tt[l] + drow) * TI + col)
        assert banks_ok(addrs)
        a[:, e] = P[addrs]
      for nt in range(4):
        mma_m16n8k8(acc[nt], a, b[nt])
    for nt in range(4):
      for e in range(4):
        addrs = []
        for l in lanes:
          j = 32 * nq + 8 * nt + 2 * tt[l] + (e & 1)
          m = m0 + gg[l] + (8 if e & 2 else 0)
          addrs.append(j * GS + m)
          G[j * GS + m] = acc[nt][l][e]
        assert banks_ok(addrs)
  S = np.array([[sf[64 - i + j] for j in range(NB)] for i in range(NB)])
  want = p_true @ S                              # [item][j]
  got = G.reshape(NB, GS)[:, :TI].T
  assert np.allclose(got, want, rtol=1e-12), np.abs(got - want).max()
  print("pa_mma index maps OK; all shared-memory accesses conflict-free")


if __name__ == "__main__":
  main()
