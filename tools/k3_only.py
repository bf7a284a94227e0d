"""Development aid: a short program for ncu captures of the masking kernel - one forward MDCT, then five ac_pa_encode_f32
launches on the bench workload (PROBE_WORKLOAD, default cfg2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200
import bench

b, c, sr, s, n = bench.workload_shape(os.environ.get("PROBE_WORKLOAD", "cfg2"))
x = bench.device_synthetic_audio(torch, b, s, c, sr, 0, torch.device("cuda"))
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
y = codec.mdct.transform(x)
for _ in range(5):
  q, step = codec.psychoacoustic.encode(y)
torch.cuda.synchronize()
print("ok", float(step.mean()))
