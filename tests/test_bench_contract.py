"""CPU suite: the part of bench.py's contract that needs no GPU -- the reference arm (`--impl reference`: the reference's CPU
algorithm, here the oracle port, timed on the host cores) prints ONE JSON line with the keys the driver reads, on the same
metric / unit / workload string as the GPU arm; and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def run_bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = run_bench("--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1", "--basis", "512")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "el+hole Chebyshev terms/s" and d["unit"] == "terms/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "N=512" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                      # on a GPU box the GPU suite covers the arm itself
    res = run_bench("--steps", "1", "--warmup", "1", "--basis", "512", "--skip-cpu", "--skip-e2e", "--skip-65k", "--skip-small")
    assert res.returncode != 0
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")], "no result line may be printed without a GPU"
