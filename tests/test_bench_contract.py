"""bench.py prints ONE JSON line carrying every key the driver's contract names (GPU box only)."""
import json
import os.path as osp
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_bench_json_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, osp.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
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
