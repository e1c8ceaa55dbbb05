"""GPU: the bench.py JSON contract (one line, required keys, roofline + cpu_baseline objects) on a short run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_bench_json_line(device):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "2", "--warmup", "3", "--no-extra"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "energy+forces atom-steps/sec" and d["unit"] == "atom-steps/s" and d["n_gpus"] == 1
    assert d["value"] > 1e5 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["config"]["workload"].startswith("C2")
    assert d["gpu_launches"] > 0
    e2e = d["e2e"]
    # positions-in leg: coordinates (f64) + atomic numbers + cells, not the graph
    assert 0 < e2e["value"] and 27648 * 24 <= e2e["h2d_bytes_per_step"] < 2e6 and e2e["d2h_bytes_per_step"] > 0
    assert abs(e2e["value"] - d["value"]) > 1e-6 * d["value"], "e2e must be measured separately"
    assert d["e2e_graph_given"]["h2d_bytes_per_step"] > 1e8
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] in ("hbm", "tensor") and 0 < rf["frac"] <= 1.0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    ck = d["clocks"]
    assert "sm_mhz" in ck and "sm_max_mhz" in ck and "reasons" in ck
