"""bench.py's reference arm (the CPU oracle port) runs without a GPU and prints the contract's JSON line; the product
arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=300)


def test_reference_arm_prints_contract_line():
    r = _run("--impl", "reference", "--workload", "smmnist_b16", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert line["config"]["workload"] == "smmnist_b16"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--workload", "smmnist_b16", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_product_arm_refuses_to_run_without_cuda():
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
