"""Development aid: one compact encode + fused compact decode on the cfg2 tensor (for an ncu capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200
import bench

b, c, sr, s, n = bench.workload_shape("cfg2")
x = bench.device_synthetic_audio(torch, b, s, c, sr, 0, torch.device("cuda"))
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
mdct, pa = codec.mdct, codec.psychoacoustic
y = mdct.transform(x)
for _ in range(3):
  q, g = pa.encode_compact(y)
  xh = mdct.inverse_transform_compact(q, g, pa)
torch.cuda.synchronize()
print("ok", float(xh.abs().max()))
