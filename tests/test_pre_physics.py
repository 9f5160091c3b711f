"""Pre-physics tensor math: action delay / noise / clip (humanoid_env.py:624-636) and the PD torque
law of `_compute_torques` (:910-925).  CPU: oracle vs the golden vector of the reference method.  GPU: kernel vs both."""
import numpy as np
import pytest
import torch

from oracle import env_oracle as eo


def _load(golden_dir):
    z = np.load(f"{golden_dir}/pre_physics_d12.npz")
    return z, (lambda k: torch.tensor(z[k]))


def test_oracle_matches_reference_torques(golden_dir):
    z, t = _load(golden_dir)
    actions = eo.mix_actions(t("raw"), t("prev"), t("delay"), t("noise"), float(z["dyn_rand"]), float(z["clip_actions"]))
    assert torch.equal(actions, t("actions"))
    got = eo.compute_torques(actions, t("dof_pos"), t("dof_vel"), t("p_gains"), t("d_gains"), t("default_dof_pos"),
                             t("torque_limits"), float(z["action_scale"]))
    assert torch.equal(got, t("torques"))


@pytest.mark.gpu
def test_action_torque_kernel(golden_dir):
    from humanoid_gym_with_comments_b200.envs import XBotEnvMath, XBotMathConfig
    z, t = _load(golden_dir)
    N, D, dev = int(z["N"]), int(z["D"]), "cuda:0"
    env = XBotEnvMath(XBotMathConfig.for_dof(D, num_envs=N), dev, default_dof_pos=t("default_dof_pos"))
    env.actions.copy_(t("prev"))
    dof_state = torch.stack((t("dof_pos"), t("dof_vel")), dim=-1).reshape(N * D, 2).to(dev).contiguous()
    torques = torch.zeros(N, D, device=dev)
    g = lambda k: t(k).to(dev).contiguous()
    acts, tq = env.pre_physics(g("raw"), dof_state, g("p_gains"), g("d_gains"), g("torque_limits"), torques,
                               delay=g("delay").reshape(-1), noise=g("noise"), dyn_rand=float(z["dyn_rand"]),
                               clip_actions=float(z["clip_actions"]), action_scale=float(z["action_scale"]))
    assert torch.equal(acts.cpu(), t("actions"))          # -fmad=false: op-for-op identical to torch
    assert torch.equal(tq.cpu(), t("torques"))
    # torque-only re-evaluation (physics sub-steps) and the Philox path
    tq2 = torch.zeros_like(torques)
    env.pre_physics(None, dof_state, g("p_gains"), g("d_gains"), g("torque_limits"), tq2)
    assert torch.equal(tq2, tq)
    env.actions.copy_(t("prev"))
    a3, _ = env.pre_physics(g("raw"), dyn_rand=0.0)
    lo, hi = torch.minimum(t("raw"), t("prev")).clip(-18, 18), torch.maximum(t("raw"), t("prev")).clip(-18, 18)
    assert torch.all(a3.cpu() >= lo - 1e-5) and torch.all(a3.cpu() <= hi + 1e-5)    # a convex mix of raw and previous
