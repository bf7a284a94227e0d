"""NumPy emulation of the thread-level algorithm of mdct_tile_kernels.cu (index maps, swizzle, tables).

Development aid only: checks the table construction and the Stockham index arithmetic against the oracle
before the CUDA version goes to a GPU.  Not part of the product, not used by the tests.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import audiocodec_oracle as oracle

PLANS = {16: (8, 8, 8, 1, 1), 32: (16, 8, 8, 2, 1), 64: (32, 8, 8, 4, 1), 128: (64, 8, 8, 8, 1),
         256: (128, 16, 16, 8, 1), 512: (256, 16, 16, 16, 1), 1024: (512, 16, 8, 8, 8),
         2048: (1024, 16, 16, 8, 8), 4096: (2048, 16, 16, 16, 8)}


class Plan:
  def __init__(self, n):
    self.M, self.E, self.R0, self.R1, self.R2 = PLANS[n]
    self.T = self.M // self.E
    self.RL = self.R2 if self.R2 > 1 else (self.R1 if self.R1 > 1 else self.R0)
    self.SH = int(np.log2(self.R0))

  def in_index(self, t, s):
    return (t + self.T * (s // self.R0)) + (s % self.R0) * (self.M // self.R0)

  def out_index(self, t, s):
    return (t + self.T * (s // self.RL)) + (s % self.RL) * (self.M // self.RL)

  def swz(self, pos):
    return pos ^ ((pos >> self.SH) & 7) if self.M >= 8 * (1 << self.SH) or True else pos


def build_tables(n, window):
  ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float64)
  h = n // 2
  F = ref.fold_matrix(np.float64)
  Finv = np.linalg.inv(F)
  p = np.arange(h)
  fold = np.stack([F[p, h - 1 - p], F[n - 1 - p, h - 1 - p], F[p, h + p], F[n - 1 - p, h + p]], 1)
  unfold = np.stack([Finv[h - 1 - p, p], Finv[h + p, p], Finv[h - 1 - p, n - 1 - p], Finv[h + p, n - 1 - p]], 1)
  M = h
  ang = -np.pi * (np.arange(M) + 0.125) / n
  c, s = np.cos(ang), np.sin(ang)
  scale_fwd, scale_inv = 1.0 / (n * np.sqrt(2.0)), 2.0 * np.sqrt(2.0)
  pre_fwd = np.zeros((2, M, 8))
  for nn in range(M):
    low = nn < M // 2
    pp = h - 1 - 2 * nn if low else 2 * nn - h
    a0, a1, a2, a3 = fold[pp]
    if low:   # z = alpha + i beta
      re = [a0 * c[nn], a1 * c[nn], -a2 * s[nn], -a3 * s[nn]]
      im = [a0 * s[nn], a1 * s[nn], a2 * c[nn], a3 * c[nn]]
    else:     # z = beta + i alpha
      re = [-a0 * s[nn], -a1 * s[nn], a2 * c[nn], a3 * c[nn]]
      im = [a0 * c[nn], a1 * c[nn], a2 * s[nn], a3 * s[nn]]
    # order of the four loads: (prev L1, prev L2, cur L1, cur L2); variant 0: L1 = x[p], L2 = x[N-1-p]
    pre_fwd[0, nn] = [re[0], re[1], re[2], re[3], im[0], im[1], im[2], im[3]]
    pre_fwd[1, nn] = [re[1], re[0], re[3], re[2], im[1], im[0], im[3], im[2]]
  post = {}
  for name, sc in (("fwd", scale_fwd), ("inv", scale_inv)):
    wx, wy = c * sc, s * sc
    tab = np.zeros((2, M, 4))
    # first store S1 = vx c0 + vy c1, second S2 = vx c2 + vy c3; variant 0: S1 -> out[2k] = Re D, S2 -> out[N-1-2k] = -Im D
    tab[0] = np.stack([wx, -wy, -wy, -wx], 1)
    tab[1] = np.stack([-wy, -wx, wx, -wy], 1)
    post[name] = tab
  pre_inv = np.zeros((2, M, 4))
  # loads L1, L2; variant 0: L1 = Y[2n], L2 = Y[N-1-2n];  Re = L1 k0 + L2 k1, Im = L1 k2 + L2 k3
  pre_inv[0] = np.stack([c, -s, s, c], 1)
  pre_inv[1] = np.stack([-s, c, c, s], 1)
  roots = np.exp(-2j * np.pi * np.arange(M) / M)
  return dict(fold=fold, unfold=unfold, pre_fwd=pre_fwd, post_fwd=post["fwd"], post_inv=post["inv"], pre_inv=pre_inv,
              roots=roots)


def dft(v):
  r = len(v)
  k = np.arange(r)
  return np.exp(-2j * np.pi * np.outer(k, k) / r) @ v


def fft_group(pl, v, roots):
  """v[t][s] complex for one group (all T threads), returns same shape in out_index order.  Emulates the exchange
  through a swizzled scratch of M slots."""
  M, E, T, R0, R1, R2 = pl.M, pl.E, pl.T, pl.R0, pl.R1, pl.R2
  for t in range(T):
    for q in range(E // R0):
      v[t][q * R0:(q + 1) * R0] = dft(v[t][q * R0:(q + 1) * R0])

  def to_buf(R, NS):
    buf = np.full(M, np.nan + 0j)
    for t in range(T):
      for q in range(E // R):
        j = t + T * q
        j0 = (j // NS) * NS * R + (j % NS)
        for r in range(R):
          ph = pl.swz(j0 + r * NS)
          assert np.isnan(buf[ph].real), "swizzle collision"
          buf[ph] = v[t][q * R + r]
    return buf

  def from_buf(buf, R, NS):
    for t in range(T):
      for q in range(E // R):
        j = t + T * q
        k = j % NS
        x = np.array([buf[pl.swz(j + r * (M // R))] for r in range(R)])
        for r in range(1, R):
          x[r] *= roots[r * k * (M // (NS * R))]
        v[t][q * R:(q + 1) * R] = dft(x)

  if R1 > 1:
    from_buf(to_buf(R0, 1), R1, R0)
    if R2 > 1:
      from_buf(to_buf(R1, R0), R2, R0 * R1)
  return v


def forward_frame(pl, n, tabs, xp, xc, variant):
  """xp, xc: previous / current block [N]; returns the N MDCT coefficients."""
  M, E, T = pl.M, pl.E, pl.T
  h = n // 2
  v = np.zeros((T, E), complex)
  for t in range(T):
    for s in range(E):
      nn = pl.in_index(t, s)
      low = nn < M // 2
      pp = h - 1 - 2 * nn if low else 2 * nn - h
      qq = n - 1 - pp
      a1, a2 = (pp, qq) if variant == 0 else (qq, pp)
      k = tabs["pre_fwd"][variant, nn]
      l = [xp[a1], xp[a2], xc[a1], xc[a2]]
      v[t, s] = complex(np.dot(k[:4], l), np.dot(k[4:], l))
  v = fft_group(pl, v, tabs["roots"])
  out = np.zeros(n)
  for t in range(T):
    for s in range(E):
      k = pl.out_index(t, s)
      c = tabs["post_fwd"][variant, k]
      s1 = v[t, s].real * c[0] + v[t, s].imag * c[1]
      s2 = v[t, s].real * c[2] + v[t, s].imag * c[3]
      i1, i2 = (2 * k, n - 1 - 2 * k) if variant == 0 else (n - 1 - 2 * k, 2 * k)
      out[i1], out[i2] = s1, s2
  return out


def inverse_frame(pl, n, tabs, y, variant):
  """y: N coefficients -> v = sqrt(4N) * DCT-IV(y)."""
  M, E, T = pl.M, pl.E, pl.T
  v = np.zeros((T, E), complex)
  for t in range(T):
    for s in range(E):
      nn = pl.in_index(t, s)
      a1, a2 = (2 * nn, n - 1 - 2 * nn) if variant == 0 else (n - 1 - 2 * nn, 2 * nn)
      k = tabs["pre_inv"][variant, nn]
      v[t, s] = complex(y[a1] * k[0] + y[a2] * k[1], y[a1] * k[2] + y[a2] * k[3])
  v = fft_group(pl, v, tabs["roots"])
  out = np.zeros(n)
  for t in range(T):
    for s in range(E):
      k = pl.out_index(t, s)
      c = tabs["post_inv"][variant, k]
      s1 = v[t, s].real * c[0] + v[t, s].imag * c[1]
      s2 = v[t, s].real * c[2] + v[t, s].imag * c[3]
      i1, i2 = (2 * k, n - 1 - 2 * k) if variant == 0 else (n - 1 - 2 * k, 2 * k)
      out[i1], out[i2] = s1, s2
  return out


def main():
  rng = np.random.default_rng(0)
  for n in (16, 32, 64, 128, 256, 512, 1024, 2048):
    for window in ("vorbis", "sine"):
      pl = Plan(n)
      tabs = build_tables(n, window)
      ref = oracle.MDCTransformer(n, window_type=window, compute_dtype=np.float64)
      blocks = 3
      x = rng.uniform(-1, 1, (1, blocks * n, 1))
      y_ref = ref.transform(x)[0, :, :, 0]
      xb = np.concatenate([np.zeros((1, n)), x[0, :, 0].reshape(blocks, n), np.zeros((1, n))])
      err = 0.0
      for f in range(blocks + 1):
        yy = forward_frame(pl, n, tabs, xb[f], xb[f + 1], f & 1)
        err = max(err, np.max(np.abs(yy - y_ref[f])))
      # inverse
      xh_ref = ref.inverse_transform(y_ref[None, :, :, None])[0, :, 0].reshape(blocks + 2, n)
      vs = [np.zeros(n)] + [inverse_frame(pl, n, tabs, y_ref[f], (f + 1) & 1) for f in range(blocks + 1)] + [np.zeros(n)]
      h = n // 2
      p = np.arange(h)
      err_i = 0.0
      for blk in range(blocks + 2):
        vn, vp = vs[blk + 1], vs[blk]
        u = tabs["unfold"]
        xo = np.zeros(n)
        xo[p] = u[:, 0] * vn[h - 1 - p] + u[:, 1] * vp[h + p]
        xo[n - 1 - p] = u[:, 2] * vn[h - 1 - p] + u[:, 3] * vp[h + p]
        err_i = max(err_i, np.max(np.abs(xo - xh_ref[blk])))
      print(n, window, "fwd err %.2e inv err %.2e" % (err, err_i))


if __name__ == "__main__":
  main()
