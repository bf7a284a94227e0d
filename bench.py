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


def run_b200_arm(args):
  import torch
  import torch.distributed as dist

  import audiocodec_b200
  from audiocodec_b200 import _capi

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device - audiocodec_b200 has no CPU path (use --impl reference for the CPU port)")
  torch.cuda.set_device(local_rank)
  device = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=device)

  b, c, sr, s, n = workload_shape(args.workload)
  if args.batch:
    b = args.batch
  frames = s // n + 1
  lib = _capi.lib()
  codec = audiocodec_b200.AudioCodec(sr, filters_n=n)
  mdct, pa = codec.mdct, codec.psychoacoustic
  mplan, pplan = mdct._plan(device), pa._plan(device)

  x = device_synthetic_audio(torch, b, s, c, sr, first_clip=rank * b, device=device)
  y = torch.empty(b, frames, n, c, device=device, dtype=torch.float32)
  q = torch.empty(b, frames, n, c, device=device, dtype=torch.int32)
  step_t = torch.empty(b, frames, n, c, device=device, dtype=torch.float32)
  xhat = torch.empty(b, (frames + 1) * n, c, device=device, dtype=torch.float32)
  stream = torch.cuda.current_stream(device)
  sp = stream.cuda_stream

  def k1():
    _capi.check(lib.ac_mdct_forward_f32(mplan, x.data_ptr(), y.data_ptr(), b, s, c, sp))

  def k3():
    _capi.check(lib.ac_pa_encode_f32(pplan, y.data_ptr(), 0.0, 1.0, step_t.data_ptr(), q.data_ptr(), b, frames, c, sp))

  def k2():
    _capi.check(lib.ac_mdct_inverse_dequant_f32(mplan, q.data_ptr(), step_t.data_ptr(), xhat.data_ptr(), b, frames, c, sp))

  kernels = [("mdct_forward", k1), ("pa_encode", k3), ("mdct_inverse_dequant", k2)]
  rows = b * c
  alg_bytes = {   # compulsory traffic only
    "mdct_forward": 4 * rows * ((frames - 1) * n + frames * n),
    "pa_encode": 4 * rows * frames * 3 * n,
    "mdct_inverse_dequant": 4 * rows * (2 * frames * n + (frames + 1) * n),
  }

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize(device)

  # ---- device-resident timing -----------------------------------------------------------------------
  for _ in range(max(args.warmup, 3)):
    for _, fn in kernels:
      fn()
  barrier()
  sampler = ClockSampler(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else
                         os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank])
  if rank == 0:
    sampler.start()
    time.sleep(0.3)
  marks = [[torch.cuda.Event(enable_timing=True) for _ in range(len(kernels) + 1)] for _ in range(args.steps)]
  launches0 = lib.ac_kernel_launch_count()
  barrier()
  t_wall0 = time.perf_counter()
  for i in range(args.steps):
    marks[i][0].record(stream)
    for j, (_, fn) in enumerate(kernels):
      fn()
      marks[i][j + 1].record(stream)
  barrier()
  t_wall1 = time.perf_counter()
  launches = lib.ac_kernel_launch_count() - launches0
  total_ms = marks[0][0].elapsed_time(marks[-1][-1])
  per_kernel_ms = {name: statistics.fmean(marks[i][j].elapsed_time(marks[i][j + 1]) for i in range(args.steps))
                   for j, (name, _) in enumerate(kernels)}

  # keep the GPU under load a little longer if the timed region was too short for a clock sample
  if rank == 0:
    t_hold = time.perf_counter()
    while time.perf_counter() - t_hold < 0.6:
      for _, fn in kernels:
        fn()
      torch.cuda.synchronize(device)
    t_load_end = time.perf_counter()

  # ---- end to end through the public API with host buffers ------------------------------------------
  e2e_ms, e2e_steps, h2d_bytes, d2h_bytes = float("nan"), 0, 0, 0
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
  if rank == 0:
    sampler.stop()

  # ---- aggregate over ranks: max time, gathered bitstream statistics ---------------------------------
  from audiocodec_b200 import sharding
  times = sharding.max_over_ranks(torch.tensor([total_ms, e2e_ms], device=device, dtype=torch.float64))
  # the one collective of the job: bitstream sizes and statistics of every shard
  stats = sharding.gather_stats(codec.stats(q).to(torch.float64)).sum(0)
  total_ms, e2e_ms = times.tolist()
  ok = bool(torch.isfinite(xhat).all().item())
  err = (xhat[:, n:-n] - x).float().pow(2).mean().sqrt().item()

  if rank == 0:
    audio_s_per_step = world * b * s / sr
    ms_per_step = total_ms / args.steps
    value = audio_s_per_step / (ms_per_step * 1e-3)
    peaks = {}
    peak_src = "fallback"
    try:
      with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
        peaks = json.load(f)
      peak_src = "measured"
    except OSError:
      pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_kernel = {}
    for name, _ in kernels:
      gbs = alg_bytes[name] / (per_kernel_ms[name] * 1e-3) / 1e9
      per_kernel[name] = {"ms": per_kernel_ms[name], "alg_bytes": alg_bytes[name], "gbs": gbs, "frac": gbs / peak}
    dominant = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
    traffic = None
    try:
      with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
        traffic = json.load(f).get(args.workload, {}).get(dominant)
    except (OSError, ValueError):
      pass
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
      "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f32", "data": "synthetic",
      "config": bench_config(args.workload, b, c, s, n),
      "e2e": {"value": audio_s_per_step / (e2e_ms * 1e-3) if e2e_steps else None, "unit": UNIT,
              "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms if e2e_steps else None,
              "steps": e2e_steps,
              "api": "AudioCodec.roundtrip_host(pinned x) -> pinned x_hat"},
      "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "kernel": dominant, "achieved": per_kernel[dominant]["gbs"], "peak": peak,
                   "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)" if peak_src == "measured" else "fallback",
                   "unit": "GB/s", "frac": per_kernel[dominant]["frac"], "traffic": traffic,
                   "traffic_source": "profiles/dram_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                     "ncu --set full capture of this kernel on this workload; not measured in this run)",
                   "chain_gbs": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9,
                   "chain_frac": sum(alg_bytes.values()) / (ms_per_step * 1e-3) / 1e9 / peak},
      "kernels": per_kernel,
      "clocks": sampler.summary(t_wall0 - 0.05, t_load_end),
      "stats": {"coefficients": stats[0].item(), "nonzero": stats[1].item(), "bits_estimate": stats[2].item(),
                "roundtrip_rms_error": err, "finite": ok},
    }
    if world == 1 and not args.no_cpu_baseline:
      v, info = cpu_port_throughput(args.workload, budget_s=20.0, steps=2, warmup=1)
      line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"]}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()
  return 0


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
  args = ap.parse_args()
  if args.impl == "reference":
    return run_reference_arm(args)
  return run_b200_arm(args)


if __name__ == "__main__":
  sys.exit(main())
