"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm runs the reference's CPU
implementation and prints one JSON line with the agreed keys; the product arm refuses to run without a device
(there is no CPU fallback to fall back to)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--queries", "50000", "--targets", "20000")   # default workload: D
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "overlap_queries_per_sec" and d["unit"] == "queries/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["vs_baseline"] is None and d["dtype"] == "u32" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("D:") and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the driver itself
    r = _run("--steps", "1", "--warmup", "1", "--queries", "1000", "--targets", "1000")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
