"""Text summaries of an Nsight Compute report for profiles/: per kernel the headline metrics, the warp-stall mix and,
from the source page, executed instructions / stall samples / shared-memory wavefronts per barrier-delimited phase.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--items N] > profiles/rNN_ncu_full_....txt

--items: (frame, channel) items per launch, to print per-item figures (220544 for cfg2).
"""
import csv
import subprocess
import sys
from collections import Counter

METRICS = [
  "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
  "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
  "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
  "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
  "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
  "lts__t_sector_hit_rate.pct", "lts__t_sector_op_read_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
  "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
  "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
  "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
  "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]


def ncu(rep, *args):
  return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), capture_output=True, text=True).stdout


def main():
  rep = sys.argv[1]
  items = int(sys.argv[sys.argv.index("--items") + 1]) if "--items" in sys.argv else 0
  rows = list(csv.reader(ncu(rep, "--page", "raw").splitlines()))
  hdr, units = rows[0], rows[1]
  for launch in rows[2:]:
    name = launch[hdr.index("Kernel Name")]
    print("----")
    print("Kernel Name =", name)
    for m in METRICS:
      if m in hdr:
        print(m, "=", launch[hdr.index(m)], units[hdr.index(m)])
    stalls = []
    for i, h in enumerate(hdr):
      if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
        try:
          stalls.append((float(launch[i]), h.split("issue_stalled_")[1].split("_per")[0]))
        except ValueError:
          pass
    print("warp stall mix (stalled warps per issue-active cycle):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    short = name.split("(")[0].split("::")[-1].split("<")[0]
    src = list(csv.reader(ncu(rep, "--page", "source", "--print-source", "sass", "--kernel-name", "regex:" + short).splitlines()))
    if len(src) < 3:
      continue
    sh = src[1]
    i_hdr = sh.index("Instructions Executed")
    body = []
    for r in src[2:]:
      if len(r) != len(sh):
        continue
      if r[i_hdr] == sh[i_hdr]:          # a second launch of the same name: the first is enough
        break
      body.append(r)
    src = src[:2] + body
    i_inst, i_smp, i_wav = sh.index("Instructions Executed"), sh.index("# Samples"), sh.index("L1 Wavefronts Shared")
    stall_cols = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
    total_smp = sum(int(r[i_smp]) for r in src[2:]) or 1
    seg, inst, smp, wav, ops, st = 0, 0, 0, 0, Counter(), Counter()
    per = f" (per item of {items})" if items else ""
    print(f"phases (between barriers){per}: executed warp instructions, share of stall samples, shared-memory wavefronts")
    div = items or 1
    for r in src[2:]:
      text = r[1].strip()
      inst += int(r[i_inst])
      smp += int(r[i_smp])
      wav += int(r[i_wav]) if r[i_wav] not in ("-", "") else 0
      op = text.split()[1] if text.startswith("@") else text.split()[0]
      ops[op.split(".")[0]] += int(r[i_inst])
      for c in stall_cols:
        if r[c] not in ("-", ""):
          st[sh[c][6:]] += int(r[c])
      if "BAR.SYNC" in text or text.startswith("EXIT") or " EXIT" in text:
        if inst / div >= 0.5:
          print(f"  phase {seg}: {inst / div:9.1f} inst  {100 * smp / total_smp:5.1f} % samples  {wav / div:7.1f} wavefronts | "
                + ", ".join(f"{k} {v / div:.1f}" for k, v in ops.most_common(10)) + " | "
                + ", ".join(f"{k} {100 * v / max(1, smp):.0f}%" for k, v in st.most_common(5)))
        seg, inst, smp, wav, ops, st = seg + 1, 0, 0, 0, Counter(), Counter()


if __name__ == "__main__":
  main()
