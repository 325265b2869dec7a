"""The parts of bench.py's contract that run without a GPU: the reference arm (the oracle port on the host cores) prints
one JSON line with the agreed keys, and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*flags):
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *flags], cwd=REPO, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_line():
    proc = _run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "instances/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["value"] > 0
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] >= 1 and base["value"] == line["value"] and base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    baseline = json.load(open(os.path.join(REPO, "BASELINE.json")))
    assert line["metric"] == baseline["metric"]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    proc = _run("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert proc.returncode != 0
    assert "no CPU fallback" in (proc.stdout + proc.stderr)
