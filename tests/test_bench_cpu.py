"""bench.py contract checks that need no GPU: the reference (CPU) arm prints one JSON line with the keys the
driver reads, and the B200 arm fails loudly instead of falling back when there is no device."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n-mc", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and abs(line["ms_per_step"] - 1000.0 / line["value"]) < 1e-6 * line["ms_per_step"]
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0 and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    from oracle import ref_runner as R
    # the unmodified reference whenever its sources are on the machine (/root/reference or the staged oracle/_ref)
    assert cb["kind"] == ("reference" if R.reference_available() else "port")
    assert cb["cores"] >= 1 and cb["value"] == line["value"] and "no extrapolation" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())      # no number from a fallback path
