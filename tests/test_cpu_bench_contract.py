"""bench.py contract checks that need no GPU: the reference arm's JSON line (the oracle port of the reference CPU
path on the host cores), its behaviour on non-zero ranks, and that our arm refuses to run without a CUDA device
(no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "32", "--samples", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert REQUIRED <= set(line), REQUIRED - set(line)
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "volumes/s"
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "slices" in cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "32", "--samples", "2"],
             env_extra={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_describe_the_same_workload():
    sys.path.insert(0, ROOT)
    try:
        import bench
    finally:
        sys.path.pop(0)
    w = bench.workload(256, 16, "trilinear")
    assert "256^3" in w and "16 z-samples" in w and "trilinear" in w
    src = open(BENCH).read()
    assert src.count('"config": config_dict(D, N, args.interp)') == 2     # ours + reference: the same dict


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_our_arm_refuses_to_run_without_cuda():
    r = _run(["--steps", "1", "--warmup", "0", "--size", "32"], timeout=300)
    assert r.returncode != 0
    assert not any(l.startswith("{") for l in r.stdout.splitlines())      # no number without the CUDA path
