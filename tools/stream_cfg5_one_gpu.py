"""Development aid / evidence: AudioCodec.roundtrip_host on HALF of cfg5 (4096 stereo clips x 30 s, 43 GB each way on the
host) through ONE GPU - the ring of chunk buffers keeps the device footprint at a few chunks (capacity for
config 5).  Pageable host memory (pinning 87 GB is not the point here); spot-checks clips against the device path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiocodec_b200

clips = int(os.environ.get("STREAM_CLIPS", "4096"))
sr, n, c = 44100, 256, 2
s = (sr * 30 // n) * n
codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
t0 = time.time()
x = torch.empty(clips, s, c, dtype=torch.float32)
g = torch.Generator().manual_seed(1)
for i in range(0, clips, 256):                      # a few distinct blocks of noise + a tone, repeated: generation is not the test
  blk = x[i:i + 256]
  if i == 0:
    blk.uniform_(-0.3, 0.3, generator=g)
    blk += 0.4 * torch.sin(torch.arange(s, dtype=torch.float32) * (2 * 3.14159265 * 440.0 / sr)).reshape(1, s, 1)
  else:
    blk.copy_(x[:blk.shape[0]])
    blk.mul_(1.0 - 0.0001 * (i // 256))
out = torch.empty(clips, s + 2 * n, c, dtype=torch.float32)
print(f"host tensors: {x.numel() * 4 / 1e9:.1f} GB in, {out.numel() * 4 / 1e9:.1f} GB out, built in {time.time() - t0:.0f} s", flush=True)
torch.cuda.synchronize()
free0, total = torch.cuda.mem_get_info()
t0 = time.time()
codec.roundtrip_host(x, out, chunk_clips=2)
dt = time.time() - t0
free1, _ = torch.cuda.mem_get_info()
print(f"roundtrip_host of {clips} clips x 30 s on one GPU: {dt:.1f} s ({clips * 30 / dt:.0f} audio-s/s from pageable host memory), "
      f"device memory taken by the pipeline {(free0 - free1) / 1e9:.2f} GB of {total / 1e9:.0f} GB")
probe = [0, clips // 2 + 1, clips - 1]
dev = torch.stack([x[i] for i in probe]).cuda()
q, step = codec.encode(dev)
ok = torch.equal(out[probe], codec.decode(q, step).cpu())
print("spot check against the device path:", "bit-identical" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
