"""bench.py's output contract on the CPU-only legs: the reference arm prints
exactly one JSON line with the keys the driver reads, and the product arm
refuses to run without a GPU instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args],
                          capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "BASELINE config 2" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm runs")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr


def test_pinned_allocation_affinity_helper_restores_the_thread():
    """bench.py's on_cpus: a no-op for None, confines the calling thread for the block and
    puts the previous affinity back; gpu_local_cpus never raises without NVML / a GPU."""
    sys.path.insert(0, ROOT)
    import bench
    import torch
    before = os.sched_getaffinity(0)
    with bench.on_cpus(None):
        assert os.sched_getaffinity(0) == before
    one = {sorted(before)[0]}
    with bench.on_cpus(one):
        assert os.sched_getaffinity(0) == one
    assert os.sched_getaffinity(0) == before
    if not torch.cuda.is_available():
        assert bench.gpu_local_cpus(torch, 0) is None
