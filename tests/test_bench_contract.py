"""bench.py's contract with the driver, as far as it can be checked without a GPU: the reference arm prints exactly one JSON
line on stdout with the agreed keys, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_agreed_keys():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1", "--tris", "5000", "--width", "160", "--height", "90",
                  "--spp", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "Msamples/s" and j["unit"] == "Msamples/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 0 and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["unit"] == "Msamples/s" and cb["sample"]
    e2e = j["e2e"]
    assert e2e == {"value": j["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank0_works():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1", "--tris", "2000", "--width", "64", "--height", "36",
                  env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0", "--tris", "2000", "--width", "64", "--height", "36", "--spp", "1")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in r.stderr or "no CPU fallback" in r.stderr
