"""Generates tests/golden/*.npz by running the UNMODIFIED reference (this container only) and checks
that the oracle restatement reproduces it.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden            # writes tests/golden/, prints the pinning report

Fixtures (all small, self-contained: inputs AND reference outputs are stored, nothing depends on
RNG reproducibility across machines):
  learner_small.npz   N=16,T=8, 705/219/12, XBot nets + hyper-parameters: rollout tensors, eps,
                      GAE returns/advantages, permutation, per-update losses, lr, parameter probes,
                      Adam state probes, after TWO consecutive reference `PPO.update()` calls.
  learner_cfg1.npz    BASELINE.json config 1 (N=64,T=24): seeds + reference outputs only
                      (inputs regenerated from the seeds by `oracle.ppo_oracle.synthetic_rollout`).
  env_d12.npz / env_d18.npz   N=16, 20 consecutive steps with resets/time-outs of the reference env
                      math under the stub isaacgym: all sim inputs, noise, and obs/priv/rew/reset.
  pre_physics_d12.npz   action delay/noise/clip (oracle) + PD torques of the reference's `_compute_torques`.
  onnx_actor_probe.npz  layer shapes + weight statistics of the reference's trained actor
                      (`humanoid/OnnxTest.onnx`), parsed with a tiny protobuf walker.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import env_oracle as eo  # noqa: E402
from oracle import ppo_oracle as po  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
XBOT_NETS = dict(actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])


def _np(t):
    return t.detach().cpu().numpy().copy()


def _ref_alg(N, T, seed):
    PPO, ActorCritic, _ = rh.load_learner()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, 12, init_noise_std=1.0, **XBOT_NETS)
    c = po.xbot_ppo_config()
    alg = PPO(ac, num_learning_epochs=c.num_learning_epochs, num_mini_batches=c.num_mini_batches,
              clip_param=c.clip_param, gamma=c.gamma, lam=c.lam, value_loss_coef=c.value_loss_coef,
              entropy_coef=c.entropy_coef, learning_rate=c.learning_rate, max_grad_norm=c.max_grad_norm,
              use_clipped_value_loss=c.use_clipped_value_loss, schedule=c.schedule, desired_kl=c.desired_kl,
              device="cpu")
    alg.init_storage(N, T, [705], [219], [12])
    return alg, c


def _oracle_params(seed):
    torch.manual_seed(seed)
    return po.init_params(705, 219, 12, XBOT_NETS["actor_hidden_dims"], XBOT_NETS["critic_hidden_dims"],
                          (128, 128), 1.0)


def _probe(t: torch.Tensor, n=64):
    flat = t.detach().reshape(-1)
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return _np(flat[idx]), np.array([float(flat.double().sum()), float(flat.double().abs().sum())])


def learner_case(N, T, seed, n_updates, store_inputs):
    """Drive reference and oracle with identical inputs; assert agreement; return the fixture dict."""
    alg, cfg = _ref_alg(N, T, seed)
    p = _oracle_params(seed)
    sd = alg.actor_critic.state_dict()
    assert list(sd.keys()) == po.state_dict_order(p), "state_dict key order"
    for k in sd:
        assert torch.equal(sd[k], p[k]), f"init mismatch {k}"
    st = po.AdamState()
    lr = cfg.learning_rate
    out = dict(N=N, T=T, seed=seed, n_updates=n_updates)
    report = {}
    for u in range(n_updates):
        gen = torch.Generator().manual_seed(1000 * seed + 17 + u)
        gen_state = gen.get_state()
        p_start = {k: v.clone() for k, v in p.items()}
        ro, last_c, inputs = po.synthetic_rollout(p, N, T, 705, 219, 12, cfg, gen, record_inputs=True)
        # replay the same stream through the reference (ppo.py:116-142)
        with torch.inference_mode():
            for t in range(T):
                # Normal.sample() consumes one normal_ of shape [N,12]; make it produce `eps`
                real_normal = torch.normal
                eps = inputs[t]["eps"]
                torch.normal = lambda mean, std, *a, **k: mean + std * eps
                try:
                    act = alg.act(ro.observations[t], ro.privileged_observations[t])
                finally:
                    torch.normal = real_normal
                assert torch.equal(act, ro.actions[t]), f"actions t={t}"
                alg.process_env_step(inputs[t]["rew"], inputs[t]["dones"], {"time_outs": inputs[t]["time_outs"]})
            alg.compute_returns(last_c)
        s = alg.storage
        for name, a, b in (("values", s.values, ro.values), ("logp", s.actions_log_prob, ro.actions_log_prob),
                           ("rewards", s.rewards, ro.rewards), ("returns", s.returns, ro.returns),
                           ("advantages", s.advantages, ro.advantages), ("mu", s.mu, ro.mu)):
            d = (a - b).abs().max().item()
            report[f"u{u}.{name}.maxabs"] = d
            assert d <= 1e-6 * max(1.0, b.abs().max().item()), (name, d)
        torch.manual_seed(4242 + u)
        perm = torch.randperm(N * T)
        torch.manual_seed(4242 + u)
        ref_losses = alg.update()
        rec = []
        o_losses = po.ppo_update(p, st, ro, perm, cfg, lr, record=rec)
        lr = o_losses[3]
        for name, a, b in (("value", ref_losses[0], o_losses[0]), ("surrogate", ref_losses[1], o_losses[1]),
                           ("lin_vel", ref_losses[3], o_losses[2])):
            report[f"u{u}.loss.{name}.rel"] = abs(a - b) / max(abs(a), 1e-12)
            assert abs(a - b) <= 2e-5 * max(abs(a), 1e-3), (name, a, b)
        assert abs(alg.learning_rate - lr) < 1e-12, (alg.learning_rate, lr)
        sd = alg.actor_critic.state_dict()
        worst = max((sd[k] - p[k]).abs().max().item() for k in sd)
        report[f"u{u}.param.maxabs"] = worst
        assert worst <= 2.5 * lr, worst      # Adam sign ambiguity on ~0 gradients is bounded by 2*lr
        # keep oracle and reference in lock-step for the next update (bitwise same start)
        for k in sd:
            p[k].copy_(sd[k])
        osd = alg.optimizer.state_dict()["state"]
        for i, k in enumerate(po.state_dict_order(p)):
            st.exp_avg[k].copy_(osd[i]["exp_avg"])
            st.exp_avg_sq[k].copy_(osd[i]["exp_avg_sq"])
        pre = f"u{u}."
        out[pre + "gen_state"] = _np(gen_state)
        out[pre + "perm"] = _np(perm).astype(np.int64)
        out[pre + "ref_losses"] = np.array([ref_losses[0], ref_losses[1], ref_losses[3]], dtype=np.float64)
        out[pre + "ref_lr"] = np.array(alg.learning_rate)
        out[pre + "oracle_minibatch_log"] = np.array(
            [[r["value"], r["surrogate"], r["lin_vel"], r["entropy"], r["kl"], r["lr"], r["grad_norm"]] for r in rec])
        out[pre + "returns"] = _np(s.returns) if store_inputs else _probe(s.returns, 256)[0]
        out[pre + "advantages"] = _np(s.advantages) if store_inputs else _probe(s.advantages, 256)[0]
        out[pre + "returns_sums"] = _probe(s.returns)[1]
        out[pre + "advantages_sums"] = _probe(s.advantages)[1]
        if store_inputs:
            for name in ("observations", "privileged_observations", "actions", "rewards", "dones", "values",
                         "actions_log_prob", "mu", "sigma"):
                out[pre + name] = _np(getattr(ro, name))
            out[pre + "last_critic_obs"] = _np(last_c)
            out[pre + "last_values"] = _np(po.value(p_start, last_c))
            out[pre + "eps"] = _np(torch.stack([i["eps"] for i in inputs]))
            out[pre + "raw_rewards"] = _np(torch.stack([i["rew"] for i in inputs]))
            out[pre + "env_dones"] = _np(torch.stack([i["dones"] for i in inputs]))
            out[pre + "time_outs"] = _np(torch.stack([i["time_outs"] for i in inputs]))
        for k in sd:
            pr, sm = _probe(sd[k])
            out[pre + "param_probe." + k] = pr
            out[pre + "param_sums." + k] = sm
        out[pre + "adam_step"] = np.array(int(osd[0]["step"]))
        out[pre + "adam_m_probe.actor.0.weight"] = _probe(osd[1]["exp_avg"])[0]
        out[pre + "adam_v_probe.actor.0.weight"] = _probe(osd[1]["exp_avg_sq"])[0]
    return out, report


def sym_case(N=16, T=8, seed=2):
    """one `PPO.update()` of the reference WITH the mirror-symmetry loss (ppo.py:93-105,197-202) against
    the oracle.  The reference builds its permutation matrices with `.cuda()` (ppo.py:96,103); in this CPU-only container
    `Tensor.cuda` is replaced by the identity while the object is constructed -- nothing else is touched."""
    import dataclasses
    PPO, ActorCritic, _ = rh.load_learner()
    obs_perm, act_perm = po.xbot_mirror_permutations()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, 12, init_noise_std=1.0, **XBOT_NETS)
    c = dataclasses.replace(po.xbot_ppo_config(), sym_loss=True, sym_coef=0.5, obs_permutation=obs_perm, act_permutation=act_perm,
                            frame_stack=15)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        alg = PPO(ac, num_learning_epochs=c.num_learning_epochs, num_mini_batches=c.num_mini_batches, clip_param=c.clip_param,
                  gamma=c.gamma, lam=c.lam, value_loss_coef=c.value_loss_coef, entropy_coef=c.entropy_coef,
                  learning_rate=c.learning_rate, max_grad_norm=c.max_grad_norm, use_clipped_value_loss=c.use_clipped_value_loss,
                  schedule=c.schedule, desired_kl=c.desired_kl, device="cpu", sym_loss=True, obs_permutation=obs_perm,
                  act_permutation=act_perm, frame_stack=15, sym_coef=c.sym_coef)
    finally:
        torch.Tensor.cuda = real_cuda
    alg.init_storage(N, T, [705], [219], [12])
    assert torch.equal(alg.obs_perm_mat, po.perm_matrix(po.stacked_obs_permutation(obs_perm, 15)))
    assert torch.equal(alg.act_perm_mat, po.perm_matrix(act_perm))
    assert float(alg.obs_perm_mat.abs().sum()) == 705 and float(alg.act_perm_mat.abs().sum()) == 12      # signed permutations
    p = _oracle_params(seed)
    gen = torch.Generator().manual_seed(77)
    ro, last_c, inputs = po.synthetic_rollout(p, N, T, 705, 219, 12, c, gen, record_inputs=True)
    with torch.inference_mode():
        for t in range(T):
            real_normal, eps = torch.normal, inputs[t]["eps"]
            torch.normal = lambda mean, std, *a, **k: mean + std * eps
            try:
                alg.act(ro.observations[t], ro.privileged_observations[t])
            finally:
                torch.normal = real_normal
            alg.process_env_step(inputs[t]["rew"], inputs[t]["dones"], {"time_outs": inputs[t]["time_outs"]})
        alg.compute_returns(last_c)
    torch.manual_seed(99)
    perm = torch.randperm(N * T)
    torch.manual_seed(99)
    ref = alg.update()
    rec = []
    st = po.AdamState()
    o = po.ppo_update(p, st, ro, perm, c, c.learning_rate, record=rec)
    ref_sym = float(ref[2])                       # the LAST minibatch's symmetry loss (ppo.py:226 returns the tensor as is)
    report = {"value.rel": abs(ref[0] - o[0]) / abs(ref[0]), "surrogate.abs": abs(ref[1] - o[1]), "lin_vel.rel": abs(ref[3] - o[2]) / abs(ref[3]),
              "sym.rel": abs(ref_sym - rec[-1]["sym"]) / abs(ref_sym), "sym": ref_sym}
    assert report["value.rel"] < 2e-5 and report["surrogate.abs"] < 2e-6 and report["lin_vel.rel"] < 2e-5 and report["sym.rel"] < 2e-5, report
    assert ref_sym > 1e-3                          # the term is live
    assert abs(alg.learning_rate - o[3]) < 1e-12
    sd = alg.actor_critic.state_dict()
    worst = max((sd[k] - p[k]).abs().max().item() for k in sd)
    assert worst <= 2.5 * o[3], worst
    report["param.maxabs"] = worst
    out = dict(N=N, T=T, seed=seed, obs_permutation=np.array(obs_perm), act_permutation=np.array(act_perm), sym_coef=np.array(c.sym_coef),
               perm=_np(perm).astype(np.int64), ref_losses=np.array([ref[0], ref[1], ref_sym, ref[3]], dtype=np.float64),
               ref_lr=np.array(alg.learning_rate),
               oracle_minibatch_log=np.array([[r["value"], r["surrogate"], r["lin_vel"], r["entropy"], r["kl"], r["lr"], r["grad_norm"], r["sym"]] for r in rec]))
    for name in ("observations", "privileged_observations", "actions", "rewards", "dones", "values", "actions_log_prob", "mu", "sigma",
                 "returns", "advantages"):
        out[name] = _np(getattr(ro, name))
    for k in sd:
        pr, sm = _probe(sd[k])
        out["param_probe." + k] = pr
        out["param_sums." + k] = sm
    return out, report


def env_case(D, N=16, steps=20, seed=3):
    spec = eo.xbot_spec(D)
    g = torch.Generator().manual_seed(seed)
    ref = rh.RefEnv(D, N, spec)
    st = eo.EnvState(spec, N)
    ep0 = torch.randint(0, 2395, (N,), generator=g)
    ep0[0], ep0[1], ep0[2], ep0[3] = 2398, 2399, 31, 63      # time-outs within the run; phase band edges
    st.episode_length_buf = ep0.clone()
    ref.e.episode_length_buf = ep0.clone()
    assert torch.equal(ref.e.noise_scale_vec, spec.noise_scale_vec())
    assert ref.e.reward_names == spec.active_rewards()
    out = dict(D=D, N=N, steps=steps, ep_len0=_np(ep0), noise_scale_vec=_np(spec.noise_scale_vec()),
               reward_names=np.array(spec.active_rewards()),
               reward_scales_dt=np.array([spec.scaled(k) for k in spec.active_rewards()], dtype=np.float64))
    keys = ("root_states", "dof_state", "contact_forces", "rigid_state", "commands", "actions", "torques",
            "rand_push_force", "rand_push_torque", "env_frictions", "body_mass")
    acc = {k: [] for k in keys + ("noise", "obs", "priv", "rew", "reset", "time_out")}
    worst = 0.0
    for s in range(steps):
        sim = eo.synthetic_sim_state(spec, N, g, upright=(s % 3 != 2), base_contact_frac=0.06)
        noise = torch.randn(N, spec.So, generator=g)
        a = ref.step(sim, noise)
        b = eo.env_step(spec, st, sim, noise)
        for k in ("obs", "priv", "rew", "reset", "time_out"):
            assert torch.equal(a[k], b[k]), (D, s, k, (a[k].float() - b[k].float()).abs().max())
            acc[k].append(_np(a[k]))
        for k in keys:
            acc[k].append(_np(sim[k]))
        acc["noise"].append(_np(noise))
    for k, v in acc.items():
        out[k] = np.stack(v)
    out["final_episode_sums"] = _np(torch.stack([ref.e.episode_sums[k] for k in spec.active_rewards()]))
    out["final_feet_air_time"] = _np(ref.e.feet_air_time)
    out["final_feet_height"] = _np(ref.e.feet_height)
    out["final_ep_len"] = _np(ref.e.episode_length_buf)
    n_reset = int(out["reset"].sum())
    n_to = int(out["time_out"].sum())
    assert n_reset > 0 and n_to > 0, (n_reset, n_to)
    return out, dict(resets=n_reset, time_outs=n_to, maxabs=worst)


def callback_case(D=12, N=32, steps=12, seed=5):
    """the reference's `_post_physics_step_callback` (+ `_resample_commands`, `_push_robots`) and the
    command resampling / episode means of `reset_idx`, run inside the stateful env loop with replayed uniform draws."""
    spec = eo.xbot_spec(D)
    cb = eo.CallbackSpec()
    g = torch.Generator().manual_seed(seed)
    ref = rh.RefEnv(D, N, spec)
    st = eo.EnvState(spec, N)
    ep0 = torch.randint(0, 2300, (N,), generator=g)
    ep0[:8] = torch.tensor([2398, 2399, 798, 799, 1598, 1599, 31, 790])     # time-outs and resampling boundaries in the run
    st.episode_length_buf = ep0.clone()
    ref.e.episode_length_buf = ep0.clone()
    counter0 = 395                                                           # common_step_counter reaches 400 (a push) at step 4
    ref.e.common_step_counter = counter0
    cmd0 = torch.cat((0.5 * torch.randn(N, 3, generator=g), 3.0 * (2 * torch.rand(N, 1, generator=g) - 1)), dim=1)
    ref.e.commands = cmd0.clone()
    st.commands, st.rand_push_force, st.rand_push_torque = cmd0.clone(), torch.zeros(N, 3), torch.zeros(N, 3)
    assert ref.e.cfg.commands.heading_command == cb.heading_command and ref.e.cfg.domain_rand.push_robots == cb.push_robots
    assert int(ref.e.cfg.commands.resampling_time / ref.e.dt) == cb.resampling_steps
    assert float(ref.e.cfg.domain_rand.push_interval) == cb.push_interval
    assert ref.e.max_episode_length_s == cb.max_episode_length_s
    for k in ("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading"):
        assert tuple(ref.e.command_ranges[k]) == getattr(cb, k), k
    keys = ("root_states", "dof_state", "contact_forces", "rigid_state", "actions", "torques", "env_frictions", "body_mass")
    outs = ("obs", "priv", "rew", "reset", "time_out", "commands", "root_states_out", "rand_push_force", "rand_push_torque")
    acc = {k: [] for k in keys + outs + ("noise", "u_cmd", "u_push", "u_cmd_reset", "episode_means", "has_means", "pushed")}
    n_resampled = 0
    for s_ in range(steps):
        sim = eo.synthetic_sim_state(spec, N, g, upright=(s_ % 3 != 2), base_contact_frac=0.06)
        noise = torch.randn(N, spec.So, generator=g)
        u_cmd, u_push, u_rst = torch.rand(N, 3, generator=g), torch.rand(N, 5, generator=g), torch.rand(N, 3, generator=g)
        a = ref.step(sim, noise, callback=dict(u_cmd=u_cmd, u_push=u_push, u_cmd_reset=u_rst))
        b = eo.env_step(spec, st, sim, noise, callback=dict(spec=cb, common_step_counter=counter0 + s_ + 1, u_cmd=u_cmd,
                                                            u_push=u_push, u_cmd_reset=u_rst))
        n_resampled += int(((ep0 + s_ + 1) % cb.resampling_steps == 0).sum()) if s_ == 0 else 0
        for k in ("obs", "priv", "rew", "reset", "time_out", "commands", "root_states", "rand_push_force", "rand_push_torque"):
            assert torch.equal(a[k], b[k]), (D, s_, k, (a[k].float() - b[k].float()).abs().max())
        assert (a["episode_means"] is None) == (b["episode_means"] is None)
        if a["episode_means"] is not None:
            assert torch.equal(a["episode_means"], b["episode_means"]), s_
        for k in ("obs", "priv", "rew", "reset", "time_out", "commands", "rand_push_force", "rand_push_torque"):
            acc[k].append(_np(a[k]))
        acc["root_states_out"].append(_np(a["root_states"]))
        acc["pushed"].append(np.array(b["pushed"]))
        acc["has_means"].append(np.array(a["episode_means"] is not None))
        acc["episode_means"].append(_np(a["episode_means"]) if a["episode_means"] is not None
                                    else np.zeros(len(spec.active_rewards()), dtype=np.float32))
        for k in keys:
            acc[k].append(_np(sim[k]))
        for k, v in (("noise", noise), ("u_cmd", u_cmd), ("u_push", u_push), ("u_cmd_reset", u_rst)):
            acc[k].append(_np(v))
    out = dict(D=D, N=N, steps=steps, ep_len0=_np(ep0), counter0=np.array(counter0), commands0=_np(cmd0))
    for k, v in acc.items():
        out[k] = np.stack(v)
    out["final_episode_sums"] = _np(torch.stack([ref.e.episode_sums[k] for k in spec.active_rewards()]))
    n_push, n_means = int(out["pushed"].sum()), int(out["has_means"].sum())
    changed = int((out["commands"][-1][:, :2] != _np(cmd0)[:, :2]).any(axis=1).sum())
    assert n_push >= 1 and n_means >= 3 and changed >= 6, (n_push, n_means, changed)
    return out, dict(pushes=n_push, steps_with_means=n_means, envs_with_new_commands=changed, resets=int(out["reset"].sum()))


# ---------------------------------------------------------------- tiny protobuf walker for the ONNX actor
def _varint(buf, i):
    v, sh = 0, 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << sh
        if not b & 0x80:
            return v, i
        sh += 7


def _fields(buf):
    i = 0
    while i < len(buf):
        key, i = _varint(buf, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = buf[i:i + ln]
            i += ln
        elif wt == 5:
            v = buf[i:i + 4]
            i += 4
        elif wt == 1:
            v = buf[i:i + 8]
            i += 8
        else:
            raise ValueError(wt)
        yield fn, wt, v


def onnx_initializers(path):
    """ModelProto.graph(7) -> GraphProto.initializer(5) -> TensorProto{dims(1), data_type(2), name(8), raw_data(9)}"""
    buf = open(path, "rb").read()
    tensors = {}
    for fn, wt, v in _fields(buf):
        if fn == 7 and wt == 2:
            for gfn, gwt, gv in _fields(v):
                if gfn == 5 and gwt == 2:
                    dims, name, raw, fdata = [], None, None, None
                    for tfn, twt, tv in _fields(gv):
                        if tfn == 1 and twt == 0:
                            dims.append(tv)
                        elif tfn == 1 and twt == 2:
                            j = 0
                            while j < len(tv):
                                d, j = _varint(tv, j)
                                dims.append(d)
                        elif tfn == 8:
                            name = bytes(tv).decode()
                        elif tfn == 9:
                            raw = bytes(tv)
                        elif tfn == 4 and twt == 2:
                            fdata = np.frombuffer(bytes(tv), dtype="<f4")
                    arr = np.frombuffer(raw, dtype="<f4") if raw is not None else fdata
                    if arr is not None and name is not None:
                        tensors[name] = arr.reshape(dims)
    return tensors


def torque_case(D=12, N=32, seed=9):
    """`_compute_torques` of the reference (humanoid_env.py:910-925) bound to synthetic tensors; the action mix
    (:624-636) is inline in `step()` and cannot run without Isaac Gym, so only its oracle restatement is stored."""
    import types
    cls, Cfg = rh.load_env_class(D)
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    e = types.SimpleNamespace()
    e.cfg = Cfg()
    e.p_gains, e.d_gains = 50 + 150 * torch.rand(N, D, generator=g), 1 + 7 * torch.rand(N, D, generator=g)
    e.default_dof_pos = 0.3 * rn(1, D)
    e.dof_pos, e.dof_vel = 0.5 * rn(N, D), 3.0 * rn(N, D)
    e.torque_limits = 20 + 80 * torch.rand(D, generator=g)
    raw, prev = 3.0 * rn(N, D), 3.0 * rn(N, D)
    delay, noise = torch.rand(N, 1, generator=g), rn(N, D)
    actions = eo.mix_actions(raw, prev, delay, noise, e.cfg.domain_rand.dynamic_randomization, e.cfg.normalization.clip_actions)
    want = cls._compute_torques(e, actions)
    got = eo.compute_torques(actions, e.dof_pos, e.dof_vel, e.p_gains, e.d_gains, e.default_dof_pos, e.torque_limits,
                             e.cfg.control.action_scale)
    assert torch.equal(want, got)
    assert float((want.abs() == e.torque_limits).float().mean()) > 0.01      # the clip is exercised
    out = dict(D=D, N=N, raw=_np(raw), prev=_np(prev), delay=_np(delay), noise=_np(noise), actions=_np(actions),
               p_gains=_np(e.p_gains), d_gains=_np(e.d_gains), default_dof_pos=_np(e.default_dof_pos), dof_pos=_np(e.dof_pos),
               dof_vel=_np(e.dof_vel), torque_limits=_np(e.torque_limits), torques=_np(want),
               dyn_rand=np.array(e.cfg.domain_rand.dynamic_randomization), clip_actions=np.array(e.cfg.normalization.clip_actions),
               action_scale=np.array(e.cfg.control.action_scale))
    return out, dict(clipped_fraction=float((want.abs() == e.torque_limits).float().mean()))


def onnx_probe():
    path = os.path.join(rh.REF_ROOT, "humanoid", "OnnxTest.onnx")
    ts = onnx_initializers(path)
    shapes = sorted((list(v.shape) for v in ts.values() if v.ndim == 2), key=lambda s: -s[0] * s[1])
    assert shapes == [[512, 705], [256, 512], [128, 256], [12, 128]], shapes
    out = {}
    for name, v in ts.items():
        out["shape." + name] = np.array(v.shape)
        out["stats." + name] = np.array([v.mean(), v.std(), np.abs(v).max()])
    return out, dict(layers=shapes)


def main():
    os.makedirs(GOLD, exist_ok=True)
    assert rh.reference_available(), "needs /root/reference"
    manifest = {"generator": "oracle/make_golden.py", "torch": torch.__version__, "files": {}}

    def save(name, d, rep):
        path = os.path.join(GOLD, name)
        np.savez_compressed(path, **d)
        manifest["files"][name] = dict(bytes=os.path.getsize(path), report=rep,
                                       sha1=hashlib.sha1(open(path, "rb").read()).hexdigest())
        print(name, os.path.getsize(path) // 1024, "KiB", json.dumps(rep)[:400])

    d, rep = learner_case(16, 8, seed=0, n_updates=2, store_inputs=True)
    save("learner_small.npz", d, rep)
    d, rep = learner_case(64, 24, seed=0, n_updates=1, store_inputs=False)
    save("learner_cfg1.npz", d, rep)
    for D in (12, 18):
        d, rep = env_case(D)
        save(f"env_d{D}.npz", d, rep)
    d, rep = torque_case()
    save("pre_physics_d12.npz", d, rep)
    d, rep = callback_case()
    save("env_callback_d12.npz", d, rep)
    d, rep = sym_case()
    save("learner_sym_small.npz", d, rep)
    d, rep = onnx_probe()
    save("onnx_actor_probe.npz", d, rep)
    json.dump(manifest, open(os.path.join(GOLD, "MANIFEST.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
