"""CPU tests of bench.py's contract pieces that need no GPU: the reference arm prints ONE JSON line with the keys the
driver reads (rank 0 only under torchrun), and `--impl ours` refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-envs", "16", "--horizon", "8"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "ppo_env_steps_per_sec" and j["unit"] == "env-steps/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["gpu_launches"] == 0
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "16 envs x 8 steps" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm's config names what it actually ran (round-1 finding: a 256-env sample was labelled 4096 envs)
    cfg = j["config"]
    assert cfg["envs_per_gpu"] == 16 == cb["envs"] and cfg["sampled"] is True and cfg["full_envs_per_gpu"] == 4096
    assert "16 envs (bounded sample of the 4096-env workload)" in cfg["workload"]


def test_reference_arm_sample_sizing():
    """Full workload when (steps + warmup) iterations fit the time budget, else the largest power-of-two sample."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    a = argparse.Namespace(steps=20, warmup=3, horizon=60, envs=4096)
    assert b._reference_arm_envs(a, rate_guess=35e3) == 2048
    assert b._reference_arm_envs(a, rate_guess=1e6) == 4096
    assert b._reference_arm_envs(a, rate_guess=10.0) == 64
    a.steps = 2
    assert b._reference_arm_envs(a, rate_guess=35e3) == 4096


def test_reference_arm_other_ranks_exit_quietly():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = run(["--steps", "1", "--no-cpu-baseline", "--no-e2e"])
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
