"""Development aid: device-resident timing of every public entry point on one workload (CUDA events)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
b, c, sr, s, n = bench.workload_shape(args.workload)
dev = torch.device("cuda")
x = bench.device_synthetic_audio(torch, b, s, c, sr, 0, dev)
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
mdct, pa = codec.mdct, codec.psychoacoustic
y = mdct.transform(x)
ton = pa.tonality(y)
thr = pa.global_masking_threshold(y, ton)
q = pa.quantize(y, thr)
E = y.numel() * 4


def timeit(name, fn, nbytes):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.reps):
    fn()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / args.reps
  print(f"{name:38s} {ms*1e3:9.1f} us  {nbytes/ms/1e6:8.0f} GB/s  {nbytes/ms/1e6/6528.4:5.2f} of measured HBM peak")


print(bench.describe(args.workload))
timeit("transform", lambda: mdct.transform(x), x.numel() * 4 + E)
timeit("inverse_transform", lambda: mdct.inverse_transform(y), 2 * E)
timeit("tonality", lambda: pa.tonality(y), E)
timeit("global_masking_threshold(y, ton)", lambda: pa.global_masking_threshold(y, ton), 2 * E)
timeit("global_masking_threshold(y, None)", lambda: pa.global_masking_threshold(y, None), 2 * E)
timeit("quantize", lambda: pa.quantize(y, thr), 3 * E)
timeit("dequantize", lambda: pa.dequantize(q, thr), 3 * E)
timeit("encode (threshold + quantise)", lambda: pa.encode(y), 3 * E)
timeit("encode, q only", lambda: pa.encode(y, return_threshold=False), 2 * E)
G = None
try:
  _, G = pa.encode_compact(y)
except NotImplementedError:
  pass
if G is not None:
  timeit("encode_compact (q + bark thresholds)", lambda: pa.encode_compact(y), 2 * E + G.numel() * 4)
  timeit("expand_threshold", lambda: pa.expand_threshold(G), E + G.numel() * 4)
  try:
    timeit("inverse_transform_compact", lambda: mdct.inverse_transform_compact(q, G, pa), 2 * E + G.numel() * 4)

    def chain_compact():
      yy = mdct.transform(x)
      qq, gg = pa.encode_compact(yy)
      return mdct.inverse_transform_compact(qq, gg, pa)

    def chain_plain():
      yy = mdct.transform(x)
      qq, st = pa.encode(yy)
      return mdct.inverse_transform_dequantized(qq, st)

    timeit("chain, per-coefficient steps", chain_plain, 8 * E)
    timeit("chain, compact side information", chain_compact, 6 * E + 2 * G.numel() * 4)
  except NotImplementedError:
    pass
timeit("inverse_transform_dequantized", lambda: mdct.inverse_transform_dequantized(q, thr), 3 * E)
timeit("add_noise", lambda: pa.add_noise(y, thr, seed=1), 3 * E)
