"""bench.py's contract on a machine without a GPU: the reference arm prints ONE JSON line with the keys the driver
reads (same metric / unit / config family as the CUDA arm, cpu_baseline describing the run, e2e with zero copies),
and the CUDA arm refuses to run instead of detouring through the CPU."""

import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=str(ROOT))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("env-steps/sec") and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    # both arms report the warm-up they honour: at least 3 (the timing rules), i.e. max(3, --warmup)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "Rough-Unitree-Go2" in d["config"]["workload"]
    sys.path.insert(0, str(ROOT))
    import bench

    # the SAME workload string as the GPU arm prints (the driver compares the two lines' config)
    assert d["config"]["workload"] == bench.workload_string(bench.TASK_DEFAULT, 4096) and "configs[2]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_cuda_arm_refuses_to_run_without_a_gpu():
    res = _run("--steps", "1", "--warmup", "3", timeout=120)
    assert res.returncode != 0
    assert "no CPU path" in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")], "no bench line may be printed"
