#!/usr/bin/env python
"""bench.py - audio-seconds per second of the encode + decode hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic clips:
    x --K1 mdct_forward--> Y --K3 pa_encode (tonality, threshold, quantise)--> (q, step)
      --K2 mdct_inverse_dequant--> x_hat
At N = 1 the workload is BASELINE.json configs[1] ("cfg2": 64 stereo clips x 10 s @ 44.1 kHz, filters_n
= 256).  Under torchrun every rank runs the same per-GPU batch on its own clips (weak scaling, no
collective on the data path); one NCCL all_gather of the per-rank bitstream statistics follows the
timed region.  Rank 0 prints ONE JSON line.

  value     audio-s/s of the whole job, inputs resident in HBM, CUDA events, max over ranks
  e2e       same metric through the public Python API (AudioCodec.encode / .decode) with pinned HOST
            buffers: H2D of x and D2H of x_hat inside the timed region, every step
  roofline  the dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration in the timed
            region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the NumPy oracle (a port: TensorFlow, which the reference needs, is not installable
            here) on this box's host cores over a bounded sample of the same workload

`--impl reference` times that CPU port alone (all host threads, bounded sample per step).
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = "audio_seconds_per_second_encode_decode"
UNIT = "audio-s/s"

# name -> (batch, channels, sample_rate, seconds, filters_n); S is truncated to a multiple of filters_n
WORKLOADS = {
  "cfg1": (1, 1, 44100, 1, 256),
  "cfg2": (64, 2, 44100, 10, 256),
  "cfg3": (256, 2, 48000, 30, 1024),
  "cfg4": (1024, 1, 44100, 10, 256),
  "cfg5shard": (1024, 2, 44100, 30, 256),    # one rank's share of cfg5 at 8 GPUs
}


def workload_shape(name):
  b, c, sr, sec, n = WORKLOADS[name]
  s = (sr * sec // n) * n
  return b, c, sr, s, n


def describe(name):
  b, c, sr, s, n = workload_shape(name)
  return (f"{name}: {b} clips x {c} ch x {s} samples @ {sr} Hz, filters_n={n}: "
          "mdct_forward -> pa_encode(tonality+threshold+quantise) -> mdct_inverse_dequant")


def bench_config(name, b, c, s, n):
  """The `config` object of the JSON line: the same for the GPU arm and the reference arm (the driver compares them)."""
  mb = 4e-6 * b * c * (s // n + 1) * n
  return {"workload": describe(name) if b == WORKLOADS[name][0] else describe(name) + f" (batch {b})",
          "per_gpu_clips": b,
          "l2": "inputs larger than L2 (x, Y, q, step, x_hat are %.0f MB each; L2 is 126 MB)" % mb,
          "timing": "GPU arm: CUDA events on the launching stream, max over ranks; reference arm: host wall clock"}


# ------------------------------------------------------------------------------------------ CPU port
def _oracle_chain(x, mdct, pa, oracle):
  y = mdct.transform(x)
  thr = pa.global_masking_threshold(y, pa.tonality(y))
  q = oracle.quantize(y, thr)
  return mdct.inverse_transform(oracle.dequantize(q, thr))


def cpu_port_throughput(name, budget_s, steps=1, warmup=0, threads=None):
  """Times the oracle chain on `threads` host threads, one clip per task.  Returns (audio-s/s, info)."""
  import numpy as np  # noqa: F401
  from concurrent.futures import ThreadPoolExecutor
  from oracle import audiocodec_oracle as oracle

  b, c, sr, s, n = workload_shape(name)
  threads = threads or os.cpu_count() or 1
  mdct = oracle.MDCTransformer(n)
  pa = oracle.PsychoacousticModel(sr, n)
  probe = oracle.synthetic_audio(1, s, c, sr)
  _oracle_chain(probe, mdct, pa, oracle)                       # warm caches / BLAS threads
  t0 = time.perf_counter()
  _oracle_chain(probe, mdct, pa, oracle)
  t_clip = time.perf_counter() - t0
  # bounded sample: as many clips per step as fit the budget over all steps, at most the batch
  per_step_budget = budget_s / max(1, steps + warmup)
  clips = int(max(1, min(b, threads * max(1, int(per_step_budget / max(t_clip, 1e-4))))))
  xs = [oracle.synthetic_audio(1, s, c, sr, first_clip=i) for i in range(clips)]
  pool = ThreadPoolExecutor(max_workers=threads)

  def one_step():
    list(pool.map(lambda x: _oracle_chain(x, mdct, pa, oracle), xs))

  # one clip per worker thread, BLAS kept single-threaded inside each: without the limit every worker's matmul
  # spawns a full BLAS team and the oversubscription costs the port a factor of three
  try:
    from threadpoolctl import threadpool_limits
    limit = threadpool_limits(limits=1)
  except ImportError:
    limit = None
  try:
    for _ in range(warmup):
      one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
      one_step()
    elapsed = time.perf_counter() - t0
  finally:
    if limit is not None:
      limit.restore_original_limits()
  pool.shutdown()
  audio_s = clips * s / sr * steps
  info = {"cores": min(threads, clips), "clips_per_step": clips, "steps": steps, "seconds": elapsed,
          "sample": f"{clips} of {b} clips of {name} per step x {steps} step(s), NumPy oracle chain "
                    f"(transform, tonality, global_masking_threshold, quantise, dequantise, inverse_transform)"}
  return audio_s / elapsed, info


def run_reference_arm(args):
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return 0
  value, info = cpu_port_throughput(args.workload, budget_s=120.0, steps=args.steps, warmup=args.warmup)
  b, c, sr, s, n = workload_shape(args.workload)
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": 1e3 * info["seconds"] / args.steps, "higher_is_better": True,
    "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": bench_config(args.workload, b, c, s, n),
    "note": "CPU port of the reference path (NumPy oracle); the reference itself needs TensorFlow, which is not "
            "installable offline",
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line))
  return 0


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
  QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
           "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
           "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.rows = []
    self.proc = None
    self.index = index

  def start(self):
    try:
      self.proc = subprocess.Popen(
        ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._pump, daemon=True)
      self.thread.start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.rows.append((time.perf_counter(), line.strip()))

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=5)
      except subprocess.TimeoutExpired:
        self.proc.kill()

  def summary(self, t0, t1):
    sm, smax, power, reasons = [], [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for t, line in self.rows:
      if t < t0 or t > t1 + 0.15:
        continue
      f = [v.strip() for v in line.split(",")]
      try:
        sm.append(float(f[0]))
        smax.append(float(f[1]))
        power.append(float(f[2]))
      except (ValueError, IndexError):
        continue
      for name, v in zip(names, f[4:8]):
        if v.lower().startswith("active"):
          reasons.add(name)
    if not sm:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
            "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm
def device_synthetic_audio(torch, b, s, c, sr, first_clip, device):
  """SURVEY.md 8(d) workload generated on the device: 0.5 sin(2 pi f_b t + phi_c) + 0.05 N(0,1), clipped."""
  g = torch.Generator(device=device).manual_seed(1234 + first_clip)
  f = 110.0 * 2.0 ** (6.0 * torch.rand(b, 1, 1, device=device, generator=g, dtype=torch.float64))
  t = torch.arange(s, device=device, dtype=torch.float64).reshape(1, s, 1)
  phi = (torch.arange(c, device=device, dtype=torch.float64) * (torch.pi / 3.0)).reshape(1, 1, c)
  x = torch.empty(b, s, c, device=device, dtype=torch.float32)
  chunk = max(1, (1 << 24) // (s * c))
  for i in range(0, b, chunk):
    j = min(b, i + chunk)
    ph = (2.0 * torch.pi / sr) * f[i:j] * t + phi
    noise = torch.randn(j - i, s, c, device=device, generator=g, dtype=torch.float32)
    x[i:j] = (0.5 * torch.sin(ph).to(torch.float32) + 0.05 * noise).clamp_(-1.0, 1.0)
  return x


class Chain:
  """The device-resident encode + decode chain of one workload on the current device: buffers, the three launches
  through the C ABI and their algorithmic bytes (compulsory traffic only)."""

  def __init__(self, torch, name, device, first_clip=0, batch=0, x=None):
    import audiocodec_b200
    from audiocodec_b200 import _capi
    self.torch, self.capi, self.lib, self.name, self.device = torch, _capi, _capi.lib(), name, device
    b, c, sr, s, n = workload_shape(name)
    if batch:
      b = batch
    self.b, self.c, self.sr, self.s, self.n = b, c, sr, s, n
    self.frames = frames = s // n + 1
    self.codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
    self.mplan, self.pplan = self.codec.mdct._plan(device), self.codec.psychoacoustic._plan(device)
    self.x = x if x is not None else device_synthetic_audio(torch, b, s, c, sr, first_clip=first_clip, device=device)
    self.y = torch.empty(b, frames, n, c, device=device, dtype=torch.float32)
    self.q = torch.empty(b, frames, n, c, device=device, dtype=torch.int32)
    self.step = torch.empty(b, frames, n, c, device=device, dtype=torch.float32)
    self.xhat = torch.empty(b, (frames + 1) * n, c, device=device, dtype=torch.float32)
    self.stream = torch.cuda.current_stream(device)
    sp = self.stream.cuda_stream
    lib, chk = self.lib, _capi.check
    self.kernels = [
      ("mdct_forward", lambda: chk(lib.ac_mdct_forward_f32(self.mplan, self.x.data_ptr(), self.y.data_ptr(), b, s, c, sp))),
      ("pa_encode", lambda: chk(lib.ac_pa_encode_f32(self.pplan, self.y.data_ptr(), 0.0, 1.0, self.step.data_ptr(),
                                                      self.q.data_ptr(), b, frames, c, sp))),
      ("mdct_inverse_dequant", lambda: chk(lib.ac_mdct_inverse_dequant_f32(self.mplan, self.q.data_ptr(), self.step.data_ptr(),
                                                                            self.xhat.data_ptr(), b, frames, c, sp))),
    ]
    rows = b * c
    self.alg_bytes = {
      "mdct_forward": 4 * rows * ((frames - 1) * n + frames * n),
      "pa_encode": 4 * rows * frames * 3 * n,
      "mdct_inverse_dequant": 4 * rows * (2 * frames * n + (frames + 1) * n),
    }
    self.audio_s = b * s / sr

  def step_once(self):
    for _, fn in self.kernels:
      fn()

  def time(self, steps, warmup, barrier):
    """`steps` passes of the chain between two CUDA events on the launching stream (nothing else is enqueued between the
    kernels: the masking kernel is launched programmatically dependent on the forward MDCT); returns the total ms."""
    torch = self.torch
    for _ in range(max(warmup, 3)):
      self.step_once()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(self.stream)
    for _ in range(steps):
      self.step_once()
    e1.record(self.stream)
    barrier()
    return e0.elapsed_time(e1)

  def split(self, steps, barrier):
    """The same `steps` passes with a CUDA event after every kernel: {kernel: mean ms} (the events between the kernels
    cost the chain ~1 % - the total of time() is the one reported)."""
    torch = self.torch
    marks = [[torch.cuda.Event(enable_timing=True) for _ in range(len(self.kernels) + 1)] for _ in range(steps)]
    barrier()
    for i in range(steps):
      marks[i][0].record(self.stream)
      for j, (_, fn) in enumerate(self.kernels):
        fn()
        marks[i][j + 1].record(self.stream)
    barrier()
    return {name: statistics.fmean(marks[i][j].elapsed_time(marks[i][j + 1]) for i in range(steps))
            for j, (name, _) in enumerate(self.kernels)}

  def kernel_table(self, per_kernel_ms, peak):
    out = {}
    for name, _ in self.kernels:
      gbs = self.alg_bytes[name] / (per_kernel_ms[name] * 1e-3) / 1e9
      out[name] = {"ms": per_kernel_ms[name], "alg_bytes": self.alg_bytes[name], "gbs": gbs, "frac": gbs / peak}
    return out


def measured_peak():
  try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
      return float(json.load(f).get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
  except (OSError, ValueError):
    return 6650.0, "fallback"


def other_workloads(torch, device, peak, barrier, skip):
  """Device-resident chain of the other BASELINE configs on this GPU (5 steps each, after the headline timing)."""
  out = {}
  for name in ("cfg1", "cfg3", "cfg4", "cfg5shard"):
    if name == skip:
      continue
    try:
      chain = Chain(torch, name, device)
      total_ms = chain.time(5, 3, barrier)
      per_kernel = chain.split(5, barrier)
      ms = total_ms / 5
      entry = {"workload": describe(name), "ms_per_step": ms, "value": chain.audio_s / (ms * 1e-3), "unit": UNIT,
               "kernels": {k: {"ms": v["ms"], "frac": v["frac"]} for k, v in chain.kernel_table(per_kernel, peak).items()}}
      if name == "cfg4":      # BASELINE configs[3]: the psychoacoustic kernel isolated (threshold + quantiser at a fixed scale)
        entry["pa_encode_only_value"] = chain.audio_s / (per_kernel["pa_encode"] * 1e-3)
      out[name] = entry
      del chain
      torch.cuda.empty_cache()
    except Exception as e:   # a workload that does not fit is reported, not fatal
      out[name] = {"error": repr(e)[:200]}
  return out


def cfg5_sweep(torch, dist, device, rank, world, barrier):
  """BASELINE configs[4] as specified: 8192 stereo clips x 30 s, filters_n = 256, sharded over the ranks (8192 / world
  clips each, contiguous slices of the batch axis, no collective on the data path), every rank working through its
  shard in chunks of at most 1024 clips (10.8 GB per tensor) over one set of buffers.  The chunk's clips are generated
  once per rank and re-used for all of its chunks (the kernels' time does not depend on the data); timed with CUDA
  events, max over ranks."""
  total_clips = 8192
  per_rank = total_clips // world
  chunk = min(1024, per_rank)
  n_chunks = per_rank // chunk
  chain = Chain(torch, "cfg5shard", device, first_clip=rank * per_rank, batch=chunk)
  for _ in range(2):
    chain.step_once()
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(chain.stream)
  for _ in range(n_chunks):
    chain.step_once()
  e1.record(chain.stream)
  barrier()
  ms = e0.elapsed_time(e1)
  from audiocodec_b200 import sharding
  ms = sharding.max_over_ranks(torch.tensor([ms], device=device, dtype=torch.float64)).item()
  audio_s = total_clips * chain.s / chain.sr
  alg = sum(chain.alg_bytes.values()) * n_chunks * world
  del chain
  torch.cuda.empty_cache()
  return {"workload": "cfg5: 8192 clips x 2 ch x 1322752 samples @ 44100 Hz, filters_n=256, %d clips per rank in %d chunk(s) of %d"
                      % (per_rank, n_chunks, chunk),
          "n_gpus": world, "ms": ms, "value": audio_s / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
          "chain_gbs_per_gpu": alg / world / (ms * 1e-3) / 1e9}


def run_b200_arm(args, stdout_fd=1):
  import torch
  import torch.distributed as dist

  from audiocodec_b200 import _capi

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device - audiocodec_b200 has no CPU path (use --impl reference for the CPU port)")
  torch.cuda.set_device(local_rank)
  device = torch.device("cuda", local_rank)
  bind_to_gpu_numa_node(local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=device)

  lib = _capi.lib()
  chain = Chain(torch, args.workload, device, first_clip=0, batch=args.batch)
  if rank:      # every rank its own clips of the batch axis
    chain.x.copy_(device_synthetic_audio(torch, chain.b, chain.s, chain.c, chain.sr, first_clip=rank * chain.b, device=device))
  b, c, sr, s, n, frames = chain.b, chain.c, chain.sr, chain.s, chain.n, chain.frames
  codec, x, q, xhat, stream = chain.codec, chain.x, chain.q, chain.xhat, chain.stream
  rows = b * c

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(device)

  # ---- device-resident timing -----------------------------------------------------------------------
  for _ in range(max(args.warmup, 3)):
    chain.step_once()
  barrier()
  sampler = ClockSampler(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                         os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
  sampler.start()
  time.sleep(0.3)
  launches0 = lib.ac_kernel_launch_count()
  t_wall0 = time.perf_counter()
  total_ms = chain.time(args.steps, 0, barrier)
  t_wall1 = time.perf_counter()
  launches = lib.ac_kernel_launch_count() - launches0 - 3 * len(chain.kernels)     # time() re-warms with three passes
  per_kernel_ms = chain.split(args.steps, barrier)

  # keep the GPU under load a little longer if the timed region was too short for a clock sample
  t_hold = time.perf_counter()
  while time.perf_counter() - t_hold < 0.6:
    chain.step_once()
    torch.cuda.synchronize(device)
  t_load_end = time.perf_counter()

  # ---- end to end through the public API with host buffers ------------------------------------------
  e2e_ms, e2e_steps, h2d_bytes, d2h_bytes, floor_ms = float("nan"), 0, 0, 0, float("nan")
  if not args.no_e2e:
    x_host = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    out_host = torch.empty(xhat.shape, dtype=torch.float32, pin_memory=True)
    h2d_bytes, d2h_bytes = x_host.numel() * 4, out_host.numel() * 4
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
      codec.roundtrip_host(x_host, out_host)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    te0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
      codec.roundtrip_host(x_host, out_host)
    e1.record(stream)
    barrier()
    te1 = time.perf_counter()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (te1 - te0)) / e2e_steps   # the call returns host data: host-side waits count
    # the floor of that call on this box with `world` ranks copying at once: the same bytes in the same chunks over
    # the same two copy streams, no kernels (all ranks run it together, max over ranks below)
    floor_ms = copy_only_floor(torch, x_host, out_host, x, xhat, barrier, e2e_steps)
  sampler.stop()

  # ---- aggregate over ranks: max time, gathered bitstream statistics ---------------------------------
  from audiocodec_b200 import sharding
  clocks = sampler.summary(t_wall0 - 0.05, t_load_end)
  times = sharding.max_over_ranks(torch.tensor([total_ms, e2e_ms, floor_ms], device=device, dtype=torch.float64))
  clock_min = -sharding.max_over_ranks(torch.tensor([-(clocks["sm_mhz"] or 0.0)], device=device, dtype=torch.float64)).item()
  # the one collective of the job: bitstream sizes and statistics of every shard
  stats = sharding.gather_stats(codec.stats(q).to(torch.float64)).sum(0)
  total_ms, e2e_ms, floor_ms = times.tolist()
  ok = bool(torch.isfinite(xhat).all().item())
  err = (xhat[:, n:-n] - x).float().pow(2).mean().sqrt().item()
  peak, peak_src = measured_peak()

  extra = {}
  if c == 2 and n == 256:
    # the single-pass encoder (x -> q, step in ONE kernel, the amplitudes never in global memory) on the
    # same tensors: reported beside the chain, used by it only if it were faster than mdct_forward + pa_encode
    sp = stream.cuda_stream

    def fused():
      _capi.check(lib.ac_codec_encode_f32(chain.mplan, chain.pplan, x.data_ptr(), 0.0, 1.0, chain.step.data_ptr(), None,
                                          q.data_ptr(), b, s, c, None, sp))
    for _ in range(3):
      fused()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(args.steps):
      fused()
    f1.record(stream)
    barrier()
    fused_ms = f0.elapsed_time(f1) / args.steps
    fused_bytes = 4 * rows * ((frames - 1) * n + 2 * frames * n)
    extra["single_pass_encoder"] = {
      "ms": fused_ms, "alg_bytes": fused_bytes, "gbs": fused_bytes / (fused_ms * 1e-3) / 1e9,
      "frac": fused_bytes / (fused_ms * 1e-3) / 1e9 / peak,
      "two_kernels_ms": per_kernel_ms["mdct_forward"] + per_kernel_ms["pa_encode"],
      "in_chain": False, "api": "ac_codec_encode_f32 (AudioCodec.encode)"}
  if not args.no_other_workloads:
    del chain.y, chain.step
    if world == 1:
      extra["other_workloads"] = other_workloads(torch, device, peak, barrier, skip=args.workload)
    try:
      extra["cfg5_sweep"] = cfg5_sweep(torch, dist, device, rank, world, barrier)
    except Exception as e:
      extra["cfg5_sweep"] = {"error": repr(e)[:200]}

  if rank == 0:
    audio_s_per_step = world * b * s / sr
    ms_per_step = total_ms / args.steps
    value = audio_s_per_step / (ms_per_step * 1e-3)
    per_kernel = chain.kernel_table(per_kernel_ms, peak)
    alg_bytes = chain.alg_bytes
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
    traffic = None
    try:
      with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
        traffic = json.load(f).get(args.workload, {}).get(dominant)
    except (OSError, ValueError):
      pass
    clocks["sm_mhz_min_over_ranks"] = clock_min
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
      "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": bench_config(args.workload, b, c, s, n),
      "e2e": {"value": audio_s_per_step / (e2e_ms * 1e-3) if e2e_steps else None, "unit": UNIT,
              "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms if e2e_steps else None,
              "steps": e2e_steps, "copy_floor_ms": floor_ms if e2e_steps else None,
              "copy_floor": "the same bytes over the same two copy streams in the same chunks, no kernels, all ranks at once, "
                            "max over ranks",
              "api": "AudioCodec.roundtrip_host(pinned x) -> pinned x_hat"},
      "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "kernel": dominant, "achieved": per_kernel[dominant]["gbs"], "peak": peak,
                   "peak_source": peak_src, "unit": "GB/s", "frac": per_kernel[dominant]["frac"], "traffic": traffic,
                   "traffic_source": "profiles/dram_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                     "ncu --set full capture of this kernel on this workload; not measured in this run)",
                   "chain_gbs": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9,
                   "chain_frac": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9 / peak},
      "kernels": per_kernel,
      "clocks": clocks,
      "stats": {"coefficients": stats[0].item(), "nonzero": stats[1].item(), "bits_estimate": stats[2].item(),
                "stream_bytes": stats[3].item(), "bits_per_coefficient": 8.0 * stats[3].item() / max(stats[0].item(), 1.0),
                "roundtrip_rms_error": err, "finite": ok},
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
      if _ORIGINAL_AFFINITY:
        os.sched_setaffinity(0, _ORIGINAL_AFFINITY)      # the CPU port gets every host core again
      v, info = cpu_port_throughput(args.workload, budget_s=20.0, steps=2, warmup=1)
      line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    sys.stdout.flush()
    os.write(stdout_fd, (json.dumps(line) + "\n").encode())
  if world > 1:
    dist.destroy_process_group()
  return 0


def copy_only_floor(torch, x_host, out_host, x_dev, out_dev, barrier, steps):
  """Pinned host -> device and device -> host copies of one step's bytes, in ~30 MB pieces on two streams at once."""
  h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
  xh, oh = x_host.view(-1), out_host.view(-1)
  xd, od = x_dev.view(-1), out_dev.view(-1)
  piece = (30 << 20) // 4

  def once():
    with torch.cuda.stream(h2d):
      for i in range(0, xh.numel(), piece):
        xd[i:i + piece].copy_(xh[i:i + piece], non_blocking=True)
    with torch.cuda.stream(d2h):
      for i in range(0, oh.numel(), piece):
        oh[i:i + piece].copy_(od[i:i + piece], non_blocking=True)

  once()
  barrier()
  t0 = time.perf_counter()
  for _ in range(steps):
    once()
  h2d.synchronize()
  d2h.synchronize()
  return 1e3 * (time.perf_counter() - t0) / steps


_ORIGINAL_AFFINITY = None


def bind_to_gpu_numa_node(local_rank):
  """Keeps this rank's host threads (and so its pinned allocations, first touch) on the CPUs nearest to its GPU: with
  eight ranks streaming at once the copies otherwise cross the socket interconnect.  Silently does nothing where the
  topology cannot be read."""
  try:
    import pynvml
    pynvml.nvmlInit()
    idx = local_rank
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
      idx = int(vis.split(",")[local_rank])
    h = pynvml.nvmlDeviceGetHandleByIndex(idx)
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    cpus = {64 * w + bit for w, m in enumerate(mask) for bit in range(64) if (m >> bit) & 1}
    global _ORIGINAL_AFFINITY
    _ORIGINAL_AFFINITY = os.sched_getaffinity(0)
    cpus &= _ORIGINAL_AFFINITY
    if cpus:
      os.sched_setaffinity(0, cpus)
  except Exception:
    pass


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
  ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
  ap.add_argument("--batch", type=int, default=0, help="override the per-GPU clip count (debug)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (exploration runs of the big configs)")
  ap.add_argument("--no-other-workloads", action="store_true", help="skip the other BASELINE configs and the cfg5 sweep")
  args = ap.parse_args()
  if args.impl == "reference":
    return run_reference_arm(args)
  # stdout carries exactly ONE line, the JSON: NCCL prints its version banner to file descriptor 1 from C when the first
  # communicator comes up, so everything before the final print goes to stderr
  sys.stdout.flush()
  saved = os.dup(1)
  os.dup2(2, 1)
  try:
    return run_b200_arm(args, stdout_fd=saved)
  finally:
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


if __name__ == "__main__":
  sys.exit(main())
