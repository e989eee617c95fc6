"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference`, the oracle's Block-WAND
port on host threads) prints one JSON line with the keys the driver reads, and the product arm refuses to run
without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ["--docs-per-segment", "20000", "--nq", "8", "--steps", "1", "--warmup", "3"]


def _bench(args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _bench(["--impl", "reference"] + TINY)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "queries/sec" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["value"] > 0
    assert d["config"]["workload"] == "or5_top100_100M_8seg"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _bench(TINY)
    assert r.returncode != 0 or not any(ln.startswith("{") for ln in r.stdout.splitlines())
    assert "NVIDIA" in r.stderr or "CUDA" in r.stderr or "no CPU fallback" in r.stderr
