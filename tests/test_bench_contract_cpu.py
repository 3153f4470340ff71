"""bench.py contract pieces that run without a GPU: the reference arm's JSON line (one line, the keys the driver reads, rank > 0
silent) and the supervised single-GPU arm failing LOUDLY -- not falling back to the CPU -- when no CUDA device is present."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True, text=True,
                          timeout=timeout)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-batch", "16"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "iisan_cached_train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "B=16" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and "workload" in d["config"]


def test_reference_arm_other_ranks_are_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "1", "--no-cpu-baseline"], env={"IISAN_BENCH_CHILD_TIMEOUT": "300"})
    assert r.returncode != 0
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert "error" in d and "value" not in d          # no number is printed by anything but the CUDA path
