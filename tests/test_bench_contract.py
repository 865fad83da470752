"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the oracle port of the reference's CPU path, the one other place outside tests/ that may execute
oracle/) prints ONE JSON line with the keys the driver reads, on the same metric / unit as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tets/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("tets/s for stable-NH grad+PSD Hessian+CSR assembly")
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) <= 1e-9 * d["value"]
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and e2e["unit"] == d["unit"]
    assert abs(e2e["value"] - d["value"]) <= 1e-9 * d["value"]
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a GPU the product arm must exit non-zero, not print a number."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0", "--newton", "0", "--no-cpu"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
