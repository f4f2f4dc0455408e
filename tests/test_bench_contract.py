"""The driver-facing contract of bench.py that can be checked without a GPU: `--impl reference` (the reference's CPU
path on this host) prints exactly ONE line on stdout, a JSON object with the keys the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold one line, got {len(lines)}: {r.stdout[:400]}"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("decode tokens/s") and d["value"] > 0 and d["n_gpus"] == 1
    assert d["steps"] == 4 and d["warmup"] == 3 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-3 * 1000
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("Llama-2-7B Q4")


def test_product_arm_refuses_to_run_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            import pytest
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout), "there must be no CPU fallback"
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")], "no result line without a GPU"
