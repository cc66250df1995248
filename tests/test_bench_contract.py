"""bench.py contract checks that need no GPU: the reference arm's JSON line and the loud failure of the product arm
without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_prints_the_contract_line():
    """`--impl reference` times the reference's own CPU path (oracle/_ref, else the C port) on a bounded sample of the
    bench workload and prints ONE JSON line with the same metric / unit / config as the product arm."""
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["unit"] == "Mcell-updates/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("Mcell-updates/sec")
    assert line["value"] > 1.0 and line["ms_per_step"] > 0
    assert line["config"]["grid"] == [1024, 1024] and "BigRoom" in line["config"]["workload"]
    # the arm describes what IT ran (a bounded sample: 128 steps, 1 listener), not the GPU arm's workload
    assert line["config"]["time_steps"] == 128 and line["config"]["sources"] == 1 and "BOUNDED SAMPLE" in line["config"]["workload"]
    assert "4000 time steps" in line["config"]["gpu_arm_workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == line["value"]
    assert "1024x1024" in cb["sample"]
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpvref_fast.so")) and " avx2" in open("/proc/cpuinfo").read():
        assert cb["fast_math"]["value"] > 1.0 and "-ffast-math" in cb["fast_math"]["flags"]      # BASELINE.md section 3, second row
    assert line["e2e"] == {"value": line["value"], "unit": "Mcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    """No CUDA device -> the product arm must not produce a number (there is no CPU solve path to fall back to)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run(["--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert p.returncode != 0
    assert not [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
