"""bench.py's output contract on the CPU: the reference arm (`--impl reference`, the reference's own C++ path through
oracle/_ref, or the oracle port) prints one JSON line with the agreed keys, and the product arm refuses to run without
a CUDA device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "c1", "--steps", "1", "--warmup", "0", "--cpu-sample", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "interpolated_queries_per_sec" and j["unit"] == "queries/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in j, k
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None and j["dtype"] == "f64"
    assert j["config"]["workload"].startswith("c1:")
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "20000" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--workload", "c1", "--steps", "1", "--warmup", "3")
    assert r.returncode != 0
    assert "no CPU path" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_under_torchrun_prints_once():
    """launched as the driver launches the N>1 arm: rank 0 alone runs and prints, the other ranks exit 0 without work"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1",
                        "--steps", "1", "--warmup", "0", "--cpu-sample", "20000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0
