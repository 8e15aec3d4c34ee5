"""bench.py's contract as far as it can be checked without a GPU: the reference arm (`--impl reference`) runs the compiled
reference on the host cores and prints ONE JSON line with the keys the driver reads; without a CUDA device the product arm
refuses to run instead of falling back to a CPU path."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line(ref):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "path-tracing throughput" and d["unit"] == "Mpaths/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert "configs[1]" in d["config"]["workload"] and d["value"] > 0 and d["ms_per_step"] > 0
    # the line says itself that the CPU leg renders a sample frame, not the 1024x1024 job
    assert d["config"]["same_config"] is False and d["config"]["sample_width"] == 128 and d["config"]["job_width"] == 1024
    assert "128x128 sample frame" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "128x128" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and 100 < d["rays_per_path"] < 130   # the reference traces zero-weight children too


def test_product_arm_fails_loudly_without_a_gpu(lib, has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{") and '"value"' in l]
