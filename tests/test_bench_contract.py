"""bench.py prints ONE JSON line carrying every key the driver's contract names."""
import json
import os.path as osp
import subprocess
import sys

import pytest

from conftest import ROOT

def test_reference_arm_prints_exactly_one_json_line_on_cpu():
    """`bench.py --impl reference` (the CPU arm): stdout carries ONE JSON line and nothing else, with the keys the
    contract names for that arm.  Runs here without a GPU (2 clouds per step)."""
    out = subprocess.run([sys.executable, osp.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-clouds", "2"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gpair/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gpair/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["B_per_step"] == 2


@pytest.mark.gpu
def test_bench_json_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, osp.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]  # stdout is reserved for the one JSON line
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "roofline_hbm", "ops"):
        assert k in d, k
    assert d["unit"] == "Gpair/s" and d["dtype"] == "f32" and d["scaling"] == "weak" and d["n_gpus"] == 1
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]  # copies are inside the e2e region
    r = d["roofline"]
    assert r["bound"] == "fp32" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    h = d["roofline_hbm"]
    assert h["bound"] == "hbm" and 0 < h["frac"] < 1.2
    assert "workload" in d["config"] and "model" not in d["config"]
    assert len(d["blocks_ms_per_step"]) >= 3 and d["spread"]["min"] <= d["ms_per_step"] <= d["spread"]["max"]
    assert d["e2e"]["d2h_bytes_per_step"] > 10_000_000  # every output is read back, not just the loss
    assert d["host_issue_ms_per_step"] < d["ms_per_step"]  # one graph launch per step
