"""CPU tests (no GPU): the oracle restatement against the golden vectors produced by the UNMODIFIED reference
(tests/golden/*.npz, generator: oracle/make_golden.py) and against independent formulas of the reference tree."""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import env_oracle as eo
from oracle import isaac_math as im
from oracle import ppo_oracle as po

NETS = dict(actor_hidden=(512, 256, 128), critic_hidden=(768, 256, 128))


def _params(seed=0):
    torch.manual_seed(seed)
    return po.init_params(705, 219, 12, NETS["actor_hidden"], NETS["critic_hidden"], (128, 128), 1.0)


def test_manifest_lists_all_fixtures(golden_dir):
    man = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))
    for name in ("learner_small.npz", "learner_cfg1.npz", "env_d12.npz", "env_d18.npz", "onnx_actor_probe.npz"):
        assert name in man["files"] and os.path.exists(os.path.join(golden_dir, name))


def test_parameter_count_and_shapes(golden_dir):
    p = _params()
    assert sum(v.numel() for v in p.values()) == 1_033_372
    z = np.load(os.path.join(golden_dir, "onnx_actor_probe.npz"))
    shapes = sorted((tuple(z[k]) for k in z.files if k.startswith("shape.") and len(z[k]) == 2), key=lambda s: -s[0] * s[1])
    assert shapes == [(512, 705), (256, 512), (128, 256), (12, 128)]  # the reference's trained actor
    assert [tuple(p[f"actor.{i}.weight"].shape) for i in (0, 2, 4, 6)] == shapes


def test_learner_small_golden(golden_dir):
    """Two consecutive reference updates: rollout tensors, GAE, losses, lr, parameters."""
    z = np.load(os.path.join(golden_dir, "learner_small.npz"))
    N, T = int(z["N"]), int(z["T"])
    cfg = po.xbot_ppo_config()
    p, st, lr = _params(int(z["seed"])), po.AdamState(), cfg.learning_rate
    t_ = torch.tensor
    for u in range(int(z["n_updates"])):
        pre = f"u{u}."
        obs, priv = t_(z[pre + "observations"]), t_(z[pre + "privileged_observations"])
        eps = t_(z[pre + "eps"])
        with torch.no_grad():
            for t in range(T):
                m = po.policy_mean(p, obs[t])
                a = po.sample_actions(m, p["std"], eps[t])
                assert torch.equal(a, t_(z[pre + "actions"][t]))
                assert torch.equal(po.value(p, priv[t]), t_(z[pre + "values"][t]))
                assert torch.equal(po.log_prob(m, p["std"], a).view(-1, 1), t_(z[pre + "actions_log_prob"][t]))
                r = po.bootstrap_rewards(t_(z[pre + "raw_rewards"][t]), t_(z[pre + "values"][t]), t_(z[pre + "time_outs"][t]), cfg.gamma)
                assert torch.equal(r.view(-1, 1), t_(z[pre + "rewards"][t]))
            last_v = po.value(p, t_(z[pre + "last_critic_obs"]))
        ret, adv = po.gae_returns(t_(z[pre + "rewards"]), t_(z[pre + "values"]), t_(z[pre + "dones"]), last_v, cfg.gamma, cfg.lam)
        assert torch.equal(ret, t_(z[pre + "returns"])) and torch.equal(adv, t_(z[pre + "advantages"]))
        ro = po.Rollout(obs, priv, t_(z[pre + "actions"]), t_(z[pre + "rewards"]), t_(z[pre + "dones"]), t_(z[pre + "values"]),
                        t_(z[pre + "actions_log_prob"]), t_(z[pre + "mu"]), t_(z[pre + "sigma"]), ret, adv)
        v, s, lv, lr = po.ppo_update(p, st, ro, t_(z[pre + "perm"]), cfg, lr)
        rv, rs, rlv = z[pre + "ref_losses"]
        assert v == pytest.approx(rv, rel=1e-6) and s == pytest.approx(rs, rel=1e-5, abs=1e-8) and lv == pytest.approx(rlv, rel=1e-6)
        assert lr == pytest.approx(float(z[pre + "ref_lr"]), rel=1e-12)
        idx = lambda t: torch.linspace(0, t.numel() - 1, min(64, t.numel())).long()
        for k in po.state_dict_order(p):
            got = p[k].reshape(-1)[idx(p[k])].numpy()
            assert np.allclose(got, z[pre + "param_probe." + k], atol=2.5 * lr), k
        assert int(z[pre + "adam_step"]) == st.step


def test_learner_cfg1_golden(golden_dir):
    """BASELINE config 1 regenerated from seeds; the reference's losses and GAE probes."""
    z = np.load(os.path.join(golden_dir, "learner_cfg1.npz"))
    N, T, seed = int(z["N"]), int(z["T"]), int(z["seed"])
    cfg = po.xbot_ppo_config()
    p = _params(seed)
    gen = torch.Generator()
    gen.set_state(torch.tensor(z["u0.gen_state"]))
    ro, _, _ = po.synthetic_rollout(p, N, T, 705, 219, 12, cfg, gen)
    sums = lambda t: np.array([float(t.double().sum()), float(t.double().abs().sum())])
    assert np.allclose(sums(ro.returns), z["u0.returns_sums"], rtol=1e-9)
    assert np.allclose(sums(ro.advantages), z["u0.advantages_sums"], rtol=1e-6, atol=1e-4)
    v, s, lv, lr = po.ppo_update(p, po.AdamState(), ro, torch.tensor(z["u0.perm"]), cfg, cfg.learning_rate)
    rv, rs, rlv = z["u0.ref_losses"]
    assert v == pytest.approx(rv, rel=1e-6) and s == pytest.approx(rs, rel=1e-5, abs=1e-8) and lv == pytest.approx(rlv, rel=1e-6)
    assert lr == pytest.approx(float(z["u0.ref_lr"]), rel=1e-12)


@pytest.mark.parametrize("D", [12, 18])
def test_env_golden(D, golden_dir):
    """20 stateful steps of the reference env math (stub isaacgym), incl. resets and time-outs: bit-exact."""
    z = np.load(os.path.join(golden_dir, f"env_d{D}.npz"))
    N, steps = int(z["N"]), int(z["steps"])
    spec = eo.xbot_spec(D)
    assert list(z["reward_names"]) == spec.active_rewards() and len(spec.active_rewards()) == 22
    assert np.array_equal(z["reward_scales_dt"], np.array([spec.scaled(k) for k in spec.active_rewards()]))
    st = eo.EnvState(spec, N)
    st.episode_length_buf = torch.tensor(z["ep_len0"])
    t_ = torch.tensor
    for s in range(steps):
        ds = t_(z["dof_state"][s])
        sim = dict(root_states=t_(z["root_states"][s]), dof_pos=ds.view(N, D, 2)[..., 0], dof_vel=ds.view(N, D, 2)[..., 1],
                   contact_forces=t_(z["contact_forces"][s]), rigid_state=t_(z["rigid_state"][s]), commands=t_(z["commands"][s]),
                   actions=t_(z["actions"][s]), torques=t_(z["torques"][s]), rand_push_force=t_(z["rand_push_force"][s]),
                   rand_push_torque=t_(z["rand_push_torque"][s]), env_frictions=t_(z["env_frictions"][s]),
                   body_mass=t_(z["body_mass"][s]))
        out = eo.env_step(spec, st, sim, t_(z["noise"][s]))
        for k in ("obs", "priv", "rew", "reset", "time_out"):
            assert torch.equal(out[k], t_(z[k][s])), (D, s, k)
    assert torch.equal(st.episode_sums, t_(z["final_episode_sums"]))
    assert torch.equal(st.episode_length_buf, t_(z["final_ep_len"]))
    assert z["reset"].sum() > 0 and z["time_out"].sum() > 0


def test_env_callback_golden(golden_dir):
    """the oracle's `_post_physics_step_callback` / `_resample_commands` / `_push_robots` / episode-mean
    restatement replays the fixture recorded from the unmodified reference bit for bit."""
    z = np.load(f"{golden_dir}/env_callback_d12.npz")
    N, steps = int(z["N"]), int(z["steps"])
    spec, cb = eo.xbot_spec(12), eo.CallbackSpec()
    st = eo.EnvState(spec, N)
    st.episode_length_buf = torch.tensor(z["ep_len0"])
    st.commands, st.rand_push_force, st.rand_push_torque = torch.tensor(z["commands0"]), torch.zeros(N, 3), torch.zeros(N, 3)
    t = lambda k, s: torch.tensor(z[k][s])
    pushes = 0
    for s in range(steps):
        sim = {k: t(k, s) for k in ("root_states", "dof_state", "contact_forces", "rigid_state", "actions", "torques",
                                    "env_frictions", "body_mass")}
        sim["dof_pos"], sim["dof_vel"] = sim["dof_state"].view(N, 12, 2)[..., 0], sim["dof_state"].view(N, 12, 2)[..., 1]
        sim["commands"] = sim["rand_push_force"] = sim["rand_push_torque"] = None      # must come from the callback state
        out = eo.env_step(spec, st, sim, t("noise", s), callback=dict(
            spec=cb, common_step_counter=int(z["counter0"]) + s + 1, u_cmd=t("u_cmd", s), u_push=t("u_push", s),
            u_cmd_reset=t("u_cmd_reset", s)))
        for k in ("obs", "priv", "rew", "reset", "time_out", "commands", "rand_push_force", "rand_push_torque"):
            assert torch.equal(out[k], t(k, s)), (s, k)
        assert torch.equal(out["root_states"], t("root_states_out", s))
        assert bool(z["pushed"][s]) == out["pushed"]
        pushes += int(out["pushed"])
        assert bool(z["has_means"][s]) == (out["episode_means"] is not None)
        if out["episode_means"] is not None:
            assert torch.equal(out["episode_means"], t("episode_means", s))
    assert pushes >= 1
    assert torch.equal(st.episode_sums, torch.tensor(z["final_episode_sums"]))


def test_learner_symmetry_loss_golden(golden_dir):
    """SURVEY 8f rank 4 (oracle side): one update WITH the mirror-symmetry loss (ppo.py:93-105,197-202) replays the
    fixture recorded from the unmodified reference (sym_loss=True, sym_coef 0.5, the test mirror permutation)."""
    import dataclasses
    z = np.load(os.path.join(golden_dir, "learner_sym_small.npz"))
    obs_perm, act_perm = po.xbot_mirror_permutations()
    assert np.array_equal(z["obs_permutation"], np.array(obs_perm)) and np.array_equal(z["act_permutation"], np.array(act_perm))
    m = po.perm_matrix(po.stacked_obs_permutation(obs_perm, 15))
    assert m.shape == (705, 705) and float(m.abs().sum()) == 705 and torch.equal(m.abs().sum(0), torch.ones(705))
    x = torch.randn(3, 705)
    assert torch.allclose((x @ m) @ m, x)                          # a mirror applied twice is the identity
    cfg = dataclasses.replace(po.xbot_ppo_config(), sym_loss=True, sym_coef=float(z["sym_coef"]), obs_permutation=obs_perm,
                              act_permutation=act_perm, frame_stack=15)
    torch.manual_seed(int(z["seed"]))
    p = po.init_params(705, 219, 12, (512, 256, 128), (768, 256, 128), (128, 128), 1.0)
    t = lambda k: torch.tensor(z[k])
    ro = po.Rollout(t("observations"), t("privileged_observations"), t("actions"), t("rewards"), t("dones"), t("values"),
                    t("actions_log_prob"), t("mu"), t("sigma"))
    ro.returns, ro.advantages = t("returns"), t("advantages")
    rec = []
    v, s, lv, lr = po.ppo_update(p, po.AdamState(), ro, t("perm"), cfg, cfg.learning_rate, record=rec)
    rv, rs, rsym, rlv = z["ref_losses"]
    assert v == pytest.approx(rv, rel=1e-6) and s == pytest.approx(rs, rel=1e-5, abs=1e-8) and lv == pytest.approx(rlv, rel=1e-6)
    assert rec[-1]["sym"] == pytest.approx(rsym, rel=1e-6) and rsym > 1e-3
    assert lr == pytest.approx(float(z["ref_lr"]), rel=1e-12)
    for k in po.state_dict_order(p):
        flat = p[k].detach().reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, min(64, flat.numel())).long()
        assert np.allclose(flat[idx].numpy(), z["param_probe." + k], rtol=0, atol=2.5 * lr), k


def test_quat_apply_heading_matches_euler_yaw():
    """`quat_apply` is restated from Isaac Gym (absent from the reference tree, parity unpinned): for a pure yaw rotation
    the heading atan2(forward_y, forward_x) must equal the yaw of `get_euler_xyz`, and the rotation must keep norms."""
    from oracle.isaac_math import euler_xyz_wrapped, quat_apply
    yaw = torch.linspace(-3.0, 3.0, 25)
    q = torch.stack((torch.zeros_like(yaw), torch.zeros_like(yaw), torch.sin(yaw / 2), torch.cos(yaw / 2)), dim=1)
    f = quat_apply(q, torch.tensor([1.0, 0.0, 0.0]).repeat(25, 1))
    assert torch.allclose(torch.atan2(f[:, 1], f[:, 0]), yaw, atol=1e-6)
    assert torch.allclose(euler_xyz_wrapped(q)[:, 2], yaw, atol=1e-6)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(64, 4, generator=g)
    q = q / q.norm(dim=1, keepdim=True)
    v = torch.randn(64, 3, generator=g)
    assert torch.allclose(quat_apply(q, v).norm(dim=1), v.norm(dim=1), atol=1e-5)
    from oracle.isaac_math import quat_rotate_inverse
    assert torch.allclose(quat_rotate_inverse(q, quat_apply(q, v)), v, atol=1e-5)


def test_obs_layout_matches_sim2sim_description():
    """Independent pin of the 47-wide frame + oldest->newest stacking (reference scripts/sim2sim.py:210-233)."""
    spec = eo.xbot_spec(12)
    N = 4
    st = eo.EnvState(spec, N)
    g = torch.Generator().manual_seed(0)
    outs = []
    for s in range(3):
        sim = eo.synthetic_sim_state(spec, N, g)
        outs.append((sim, eo.env_step(spec, st, sim, None)))
    sim, out = outs[-1]
    frame = out["obs"][:, -47:]
    assert torch.allclose(frame[:, 2:5], sim["commands"][:, :3] * torch.tensor([2.0, 2.0, 1.0]))
    assert torch.allclose(frame[:, 5:17], sim["dof_pos"]) and torch.allclose(frame[:, 17:29], sim["dof_vel"] * 0.05)
    assert torch.allclose(frame[:, 41:44], out["base_ang_vel"]) and torch.allclose(frame[:, 44:47], out["euler"])
    assert torch.equal(out["obs"][:, -94:-47], outs[-2][1]["obs"][:, -47:])          # previous frame one slot older
    assert torch.equal(out["priv"][:, 2 * 73 + 53:2 * 73 + 56], (out["base_lin_vel"] * 2).clip(-18, 18))
    assert torch.all(out["obs"][:, :47 * 12] == 0)                                   # 12 oldest frames still empty


def test_isaac_math_vs_sim2sim_formulas():
    """Cross-check of the restated Isaac Gym helpers (parity unpinned upstream) against the numpy formulas the
    reference itself ships in scripts/sim2sim.py:57-77 (euler) and :87-89 (inverse rotation)."""
    g = torch.Generator().manual_seed(1)
    q = torch.randn(256, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    v = torch.randn(256, 3, generator=g, dtype=torch.float64)
    x, y, z, w = q.T
    roll = torch.atan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    pitch = torch.asin(torch.clip(2 * (w * y - z * x), -1, 1))
    yaw = torch.atan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    want = torch.stack((roll, pitch, yaw), 1)
    got = im.euler_xyz_wrapped(q)
    assert torch.allclose(got, want, atol=1e-9)
    # rotation matrix of q, applied inversely (R^T v)
    R = torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], 1),
                     torch.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], 1),
                     torch.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)], 1)
    assert torch.allclose(im.quat_rotate_inverse(q, v), torch.einsum("nji,nj->ni", R, v), atol=1e-12)


def test_gae_properties():
    """Size-independent properties: linearity in rewards at fixed values/dones; done cuts the recursion."""
    g = torch.Generator().manual_seed(2)
    T, N = 13, 7
    val, last = torch.zeros(T, N, 1), torch.zeros(N, 1)
    dones = (torch.rand(T, N, 1, generator=g) < 0.2).to(torch.uint8)
    r1, r2 = torch.rand(T, N, 1, generator=g), torch.rand(T, N, 1, generator=g)
    f = lambda r: po.gae_returns(r, val, dones, last, 0.99, 0.95)[0]
    assert torch.allclose(f(r1) + f(r2), f(r1 + r2), atol=1e-5)
    ret = f(r1)
    t_done = dones[:, 0, 0].nonzero()
    if len(t_done):
        t = int(t_done[0])
        assert ret[t, 0, 0] == r1[t, 0, 0]


def test_adaptive_lr_rule():
    assert po.adaptive_lr(1e-3, 0.03, 0.01) == pytest.approx(1e-3 / 1.5)
    assert po.adaptive_lr(1e-3, 0.001, 0.01) == pytest.approx(1.5e-3)
    assert po.adaptive_lr(1e-3, 0.0, 0.01) == 1e-3 and po.adaptive_lr(1e-3, 0.01, 0.01) == 1e-3
    assert po.adaptive_lr(1.1e-5, 1.0, 0.01) == 1e-5 and po.adaptive_lr(9e-3, 1e-4, 0.01) == 1e-2
