"""bench.py contract on the CPU side: the reference arm (the oracle timed on the host cores)
prints one JSON line with the keys the driver reads, and needs neither a GPU nor the CUDA
library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(oracle):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "tiny", "--steps", "2", "--warmup", "1", "--ref-sample-sweeps", "40"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spin_flip_attempts_per_sec" and d["unit"] == "flips/s"
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["scaling"] == "strong" and isinstance(d["dtype"], str) and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


import pytest


@pytest.mark.gpu
def test_b200_arm_prints_the_contract_line(native):
    """The product arm on the CI-size workload: one JSON line with value / e2e / roofline /
    clocks / gpu_launches, all measured (no CPU baseline leg here, it has its own test)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "3",
                          "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert "impl" not in d or d["impl"] != "reference"
    assert d["metric"] == "spin_flip_attempts_per_sec" and d["unit"] == "flips/s" and d["n_gpus"] == 1
    assert d["steps"] == 3 and d["warmup"] == 3 and d["value"] > 0 and d["higher_is_better"] is True
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e2e = d["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] >= 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] <= d["value"] * 1.05      # host copies inside the timed region cannot make it faster
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]
