"""CPU: the benchmark's reference arm runs without a GPU and prints the contract's JSON line; the CUDA arm
refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(args, **env):
    e = dict(os.environ, **env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=600)


def test_reference_arm_json_line():
    p = _run(["--impl", "reference", "--config", "c3", "--steps", "1", "--warmup", "0", "--cpu-sample-stride", "16"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ray_steps_per_s" and d["unit"] == "ray-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config3" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--config", "c3", "--gpus", "2", "--steps", "1", "--warmup", "0"], RANK="1",
             WORLD_SIZE="2", LOCAL_RANK="1")
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    p = _run(["--config", "c3", "--steps", "1", "--warmup", "3"])
    assert p.returncode != 0
    assert "CUDA" in (p.stderr + p.stdout)


def test_reference_numpy_leg_runs_the_unmodified_reference():
    """bench.py's `extras.reference_numpy`: the UNMODIFIED reference package (baseline/_ref, installed by
    baseline/install_ref.sh) timed in a child process: ray_trace single-process and over a process pool, the CPU sampler."""
    import pytest
    sys.path.insert(0, str(ROOT))
    import bench
    if not (ROOT / "baseline" / "_ref" / "raytracingGRFF").is_dir():
        pytest.skip("baseline/_ref is not installed here")
    w = bench.workload("c3", 1)
    w["grid_n"] = 32
    d = bench.reference_numpy_timing(None, w, budget_rays=64, n_steps=12)
    assert "unavailable" not in d, d
    assert d["rays"] == 64 and d["ray_trace_single_process_ray_steps_per_s"] > 0
    assert d["ray_trace_process_pool_ray_steps_per_s"] > 0 and d["sampler_samples_per_s"] > 0
