"""The UNMODIFIED reference `OnPolicyRunner` (humanoid/algo/ppo/on_policy_runner.py:47-182,293-310), loaded from
/root/reference or its byte-identical copy oracle/_ref, driven end to end on CPU with the reference's own PPO and the
reference's own env tensor math (stub isaacgym): proves that the harness used by `bench.py --impl reference` and by
tests/test_runner_gpu.py runs the real loop -- rollout, compute_returns, update, log, save, load."""
import os

import pytest
import torch

from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree (or oracle/_ref) not present")


def test_unmodified_runner_learns_saves_and_loads_on_cpu(tmp_path, monkeypatch):
    from oracle import ref_vecenv as rv
    monkeypatch.chdir(tmp_path)                      # wandb (offline mode, as the runner asks) writes ./wandb
    monkeypatch.setenv("WANDB_SILENT", "true")
    opr = rh.load_runner()
    torch.manual_seed(0)
    env = rv.RefSyntheticVecEnv(16)
    cfg = rv.train_cfg_dict(num_steps_per_env=4, save_interval=1)
    runner = opr.OnPolicyRunner(env, cfg, log_dir=str(tmp_path / "run"), device="cpu")
    assert type(runner.alg).__module__ == "humanoid.algo.ppo.ppo"
    runner.learn(2)
    assert runner.current_learning_iteration == 2
    assert {"model_0.pt", "model_1.pt", "model_2.pt"} <= set(os.listdir(tmp_path / "run"))
    w = runner.alg.actor_critic.actor[0].weight.clone()
    with torch.no_grad():
        runner.alg.actor_critic.actor[0].weight.zero_()
    runner.load(str(tmp_path / "run" / "model_2.pt"))
    assert torch.equal(runner.alg.actor_critic.actor[0].weight, w)
    assert env.course_gain == 1.0
