"""CPU checks of bench.py's contract: workload table, the reference arm's JSON line, refusal without a GPU."""

import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workload_shapes_match_survey():
  # cfg1 S = 44032, cfg2 S = 440832, cfg3 S = 1439744, cfg5 S = 1322752
  assert bench.workload_shape("cfg1") == (1, 1, 44100, 44032, 256)
  assert bench.workload_shape("cfg2") == (64, 2, 44100, 440832, 256)
  assert bench.workload_shape("cfg3") == (256, 2, 48000, 1439744, 1024)
  assert bench.workload_shape("cfg4") == (1024, 1, 44100, 440832, 256)
  assert bench.workload_shape("cfg5shard")[3] == 1322752


def test_reference_arm_prints_one_contract_line():
  proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                         "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert proc.returncode == 0, proc.stderr[-2000:]
  lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
  assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
  assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
  env = dict(os.environ, RANK="1", WORLD_SIZE="2")
  proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert proc.returncode == 0 and proc.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_b200_arm_refuses_to_run_without_a_gpu():
  proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=300, cwd=ROOT)
  assert proc.returncode != 0
  assert "no CPU path" in (proc.stderr + proc.stdout)
