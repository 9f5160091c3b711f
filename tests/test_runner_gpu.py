"""The UNMODIFIED reference `OnPolicyRunner` driving the B200 classes (-m gpu): bound exactly as INTEGRATION.md 2(a) says
(`opr.PPO, opr.ActorCritic = ...` -- the runner builds both with eval() in its own module globals,
on_policy_runner.py:66,70), `learn()` for two iterations on the synthetic VecEnv, then `save` / `load`
(on_policy_runner.py:293-310).  The runner module comes from oracle/_ref (byte-identical copy, recipe oracle/build_ref.py)."""
import os

import pytest
import torch

from gpu_util import DEV
from oracle import ref_harness as rh

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rh.reference_available(), reason="reference tree (or oracle/_ref) not present")]


@pytest.mark.parametrize("bound", [False, True])
def test_unmodified_runner_drives_b200_ppo(tmp_path, monkeypatch, bound):
    from humanoid_gym_with_comments_b200 import _lib
    from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
    from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig
    from oracle import ref_vecenv as rv
    monkeypatch.chdir(tmp_path)
    monkeypatch.setenv("WANDB_SILENT", "true")
    opr = rh.load_runner()
    monkeypatch.setattr(opr, "PPO", PPO)
    monkeypatch.setattr(opr, "ActorCritic", ActorCritic)
    N, T = 512, 8
    torch.manual_seed(0)
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), DEV)
    cfg = rv.train_cfg_dict(num_steps_per_env=T, save_interval=1)
    runner = opr.OnPolicyRunner(env, cfg, log_dir=str(tmp_path / "run"), device=DEV)
    assert isinstance(runner.alg, PPO) and isinstance(runner.alg.actor_critic, ActorCritic)
    assert runner.alg.storage.num_transitions_per_env == T and runner.alg.gamma == 0.994 and runner.alg.schedule == "adaptive"
    if bound:                                                  # optional fast path
        env.bind_storage(runner.alg.storage, gamma=runner.alg.gamma)
    l0 = _lib.load().hg_launch_count()
    runner.learn(2)
    assert _lib.load().hg_launch_count() - l0 > 2 * T          # the CUDA library did the work
    assert runner.current_learning_iteration == 2 and runner.tot_timesteps == 2 * N * T
    assert 1e-5 <= runner.alg.learning_rate <= 1e-2
    files = set(os.listdir(tmp_path / "run"))
    assert {"model_0.pt", "model_1.pt", "model_2.pt"} <= files and any(f.startswith("events.out.tfevents") for f in files)
    # save -> perturb -> load restores parameters, Adam state and the iteration counter
    ac = runner.alg.actor_critic
    obs = env.get_observations().clone()
    want = ac.act_inference(obs).clone()
    m_before = ac.flat.m.clone()
    step_before = float(runner.alg.optimizer.state_dict()["state"][0]["step"])
    assert step_before == 2 * 8
    with torch.no_grad():
        for p in ac.parameters():
            p.add_(0.01)
        ac.flat.m.zero_()
    ac.flat.refresh_shadows()
    assert not torch.allclose(ac.act_inference(obs), want)
    runner.current_learning_iteration = 0
    runner.load(str(tmp_path / "run" / "model_2.pt"))
    assert runner.current_learning_iteration == 2
    assert torch.equal(ac.act_inference(obs), want)
    assert torch.equal(ac.flat.m, m_before)
    assert float(runner.alg.optimizer.state_dict()["state"][0]["step"]) == step_before
    # and training continues from the loaded state
    runner.learn(1)
    assert runner.current_learning_iteration == 3
    policy = runner.get_inference_policy(device=DEV)
    assert policy(obs).shape == (N, 12)
