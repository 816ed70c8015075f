"""bench.py prints ONE JSON line with the keys the driver reads (contract in the task statement):
the reference arm on the CPU, and -- on a GPU box -- the B200 arm with roofline / cpu_baseline / e2e."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] in base["metric"] and d["unit"] == "MS/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run("--steps", "3", "--warmup", "3", "--seconds", "0.5", "--no-others")   # the other configs are the driver run's business
    assert BASE_KEYS <= set(d) and d.get("impl") != "reference"
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["n_gpus"] == 1
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 2 * e["h2d_bytes_per_step"]
    assert e["value"] < d["value"]                       # host copies are inside the timed region
    assert d["gpu_launches"] == d["steps"] and "sm_mhz" in d["clocks"]
    assert e["matches_device_path"] is True and 0 < e["frac_of_copy_ceiling"] < 1.5
    p = d["parity"]["configs[1]"]
    assert p["ok"] and p["rows_checked"] >= 32 and p["max_rel_rms"] <= 1e-5
    assert d["e2e_pdw"]["value"] > 0 and 0 < r["frac_sustained"] < 1.2
    n = r["same_traffic_noop"]                           # K1 alone over the same buffers: what this traffic mix reaches
    assert 0 < n["frac_of_peak"] < 1.2 and n["ms_per_launch"] > 0 and abs(n["this_kernel_vs_noop"] * r["ms_per_launch"] - n["ms_per_launch"]) < 1e-9
