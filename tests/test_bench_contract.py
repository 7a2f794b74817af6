"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on
stdout with the keys the driver reads, ranks other than 0 print nothing, and the product arm
refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    e.setdefault("OMP_NUM_THREADS", "4")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = run(["--impl", "reference", "--steps", "6", "--warmup", "1", "--workload", "test_c3_480x270"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference"
    assert j["metric"] == "4K HDR10->SDR EWA-Lanczos frames/sec" and j["unit"] == "frames/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["n_gpus"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["steps"] >= 1
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert j["config"]["workload"] == "test_c3_480x270" and j["config"]["src"] == [480, 270]
    assert j["steps"] >= 5 and "median" in cb["sample"]


def test_reference_arm_other_ranks_are_silent():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(l.strip().startswith("{") and '"value"' in l for l in r.stdout.splitlines())
