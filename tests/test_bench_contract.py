"""bench.py contract on the CPU: the reference arm prints ONE JSON line with the keys the driver reads, for the headline
workload and for the hot_plate1 (EKLT) workload; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=300):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, env=env, cwd=ROOT)


def check_reference_line(line, metric_prefix, unit):
    assert line["impl"] == "reference"
    assert line["metric"].startswith(metric_prefix) and line["unit"] == unit
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["gpu_launches"] == 0 and line["data"] == "synthetic"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_reference_arm_headline_workload():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-events", "65536")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    check_reference_line(json.loads(lines[0]), "events/s fwd+bwd", "events/s")


def test_reference_arm_eklt_workload():
    r = run_bench("--impl", "reference", "--workload", "eklt", "--steps", "2", "--warmup", "0", "--solve-events", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    check_reference_line(line, "windows/s hot_plate1 EKLT", "windows/s")
    assert line["dtype"] == "f64"


def test_reference_arm_other_ranks_stay_silent():
    env_rank = dict(os.environ, RANK="1", PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env_rank, cwd=ROOT)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a device")
def test_product_arm_needs_a_device():
    r = run_bench("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
