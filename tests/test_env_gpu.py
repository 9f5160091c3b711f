"""Fused env-step kernel vs the golden vectors produced by the UNMODIFIED reference env math and vs the oracle (-m gpu).

Tolerances: reset / time_out flags and history indexing (which frame lands where, what is zeroed) are exact;
fp32 observations / rewards agree within 1e-5 relative (+1e-6 absolute for values near zero: CUDA's sinf/cosf/expf/
atan2f differ from the CPU libm by <= 2 ulp)."""
import numpy as np
import pytest
import torch

from gpu_util import DEV
from humanoid_gym_with_comments_b200._lib import ENV_PHASE_OBS, ENV_PHASE_REWARD
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotEnvMath, XBotMathConfig, synthetic_sim_state
from oracle import env_oracle as eo

pytestmark = pytest.mark.gpu
SIM_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_state", "commands", "torques", "rand_push_force",
            "rand_push_torque", "env_frictions", "body_mass")


def close(a, b, rtol=1e-5, atol=1e-6):
    return torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())


@pytest.mark.parametrize("D", [12, 18])
def test_env_golden_replay(D, golden_dir):
    z = np.load(f"{golden_dir}/env_d{D}.npz")
    N, steps = int(z["N"]), int(z["steps"])
    cfg = XBotMathConfig.for_dof(D, num_envs=N)
    assert cfg.num_observations == z["obs"].shape[-1] and cfg.num_privileged_obs == z["priv"].shape[-1]
    env = XBotEnvMath(cfg, DEV)
    assert torch.equal(env.noise_scale_vec.cpu(), torch.tensor(z["noise_scale_vec"]))
    env.episode_length_buf.copy_(torch.tensor(z["ep_len0"]))
    env.reward_terms = torch.zeros(22, N, device=DEV)
    for s in range(steps):
        sim = {k: torch.tensor(z[k][s]).to(DEV).contiguous() for k in SIM_KEYS}
        sim["actions"] = torch.tensor(z["actions"][s]).to(DEV)
        noise = torch.tensor(z["noise"][s]).to(DEV)
        obs, priv, rew, reset, tout = env.step(sim, noise)
        torch.cuda.synchronize()
        assert torch.equal(reset.cpu(), torch.tensor(z["reset"][s])), f"step {s}: reset flags"
        assert torch.equal(tout.cpu(), torch.tensor(z["time_out"][s])), f"step {s}: time-out flags"
        for name, got, want in (("obs", obs, z["obs"][s]), ("priv", priv, z["priv"][s]), ("rew", rew, z["rew"][s])):
            ok, err = close(got.cpu(), torch.tensor(want))
            assert ok, f"D={D} step {s} {name}: max abs err {err}"
        # history indexing is exact: wherever the reference holds an exact zero (reset-zeroed frames), so do we
        zero_ref = torch.tensor(z["obs"][s]) == 0
        assert torch.equal((obs.cpu() == 0) | ~zero_ref, torch.ones_like(zero_ref)), f"step {s}: zeroed history frames"
    ok, err = close(env.episode_sums_tensor.cpu(), torch.tensor(z["final_episode_sums"]), rtol=1e-5, atol=1e-6)
    assert ok, err
    assert torch.equal(env.episode_length_buf.cpu(), torch.tensor(z["final_ep_len"]))
    assert close(env.feet_air_time.cpu(), torch.tensor(z["final_feet_air_time"]))[0]
    assert close(env.feet_height.cpu(), torch.tensor(z["final_feet_height"]))[0]


@pytest.mark.parametrize("D", [12, 18])
def test_env_vs_oracle_4096(D):
    """BASELINE config 3: 4096 envs, 20 consecutive steps with resets, against the oracle."""
    N, steps = 4096, 20
    spec = eo.xbot_spec(D)
    cfg = XBotMathConfig.for_dof(D, num_envs=N)
    g = torch.Generator().manual_seed(11)
    st = eo.EnvState(spec, N)
    st.episode_length_buf = torch.randint(0, 2400, (N,), generator=g)
    env = XBotEnvMath(cfg, DEV)
    env.episode_length_buf.copy_(st.episode_length_buf)
    worst = dict(obs=0.0, priv=0.0, rew=0.0)
    for s in range(steps):
        sim = eo.synthetic_sim_state(spec, N, g, upright=(s % 2 == 0), base_contact_frac=0.02)
        noise = torch.randn(N, spec.So, generator=g)
        want = eo.env_step(spec, st, sim, noise)
        dsim = {k: sim[k].to(DEV).contiguous() for k in SIM_KEYS}
        dsim["actions"] = sim["actions"].to(DEV)
        obs, priv, rew, reset, tout = env.step(dsim, noise.to(DEV))
        torch.cuda.synchronize()
        assert torch.equal(reset.cpu(), want["reset"]) and torch.equal(tout.cpu(), want["time_out"]), f"step {s} flags"
        for name, got in (("obs", obs), ("priv", priv), ("rew", rew)):
            ok, err = close(got.cpu(), want[name])
            worst[name] = max(worst[name], err)
            assert ok, f"D={D} step {s} {name}: {err}"
    print("worst abs err", worst)


@pytest.mark.parametrize("N,D", [(1, 12), (17, 12), (1003, 12), (1003, 18)])
def test_env_ragged_env_counts(N, D):
    """Env counts that are not a multiple of the 16-env block (ragged last block, a single env), with padded
    destination rows (the learner's storage layout: fp32 row pitch = width, bf16 shadow pitch = 768 / 256 or 1024 / 320)."""
    spec = eo.xbot_spec(D)
    cfg = XBotMathConfig.for_dof(D, num_envs=N)
    g = torch.Generator().manual_seed(5)
    st = eo.EnvState(spec, N)
    st.episode_length_buf = torch.randint(0, 2400, (N,), generator=g)
    env = XBotEnvMath(cfg, DEV)
    env.episode_length_buf.copy_(st.episode_length_buf)
    ld_o, ld_p = (spec.num_obs + 63) // 64 * 64, (spec.num_priv + 63) // 64 * 64
    for s in range(4):
        sim = eo.synthetic_sim_state(spec, N, g, base_contact_frac=0.1)
        noise = torch.randn(N, spec.So, generator=g)
        want = eo.env_step(spec, st, sim, noise)
        dsim = {k: sim[k].to(DEV).contiguous() for k in SIM_KEYS}
        dsim["actions"] = sim["actions"].to(DEV)
        o16 = torch.full((N, ld_o), 7.0, dtype=torch.bfloat16, device=DEV)
        p16 = torch.full((N, ld_p), 7.0, dtype=torch.bfloat16, device=DEV)
        obs, priv, rew, reset, tout = env.step(dsim, noise.to(DEV), obs16_out=o16, priv16_out=p16)
        assert torch.equal(reset.cpu(), want["reset"]) and torch.equal(tout.cpu(), want["time_out"])
        for name, got in (("obs", obs), ("priv", priv), ("rew", rew)):
            ok, err = close(got.cpu(), want[name])
            assert ok, f"N={N} D={D} step {s} {name}: {err}"
        assert torch.equal(o16[:, :spec.num_obs].float(), obs.to(torch.bfloat16).float()), "bf16 shadow of obs"
        assert torch.equal(p16[:, :spec.num_priv].float(), priv.to(torch.bfloat16).float()), "bf16 shadow of priv"
        assert (o16[:, spec.num_obs:] == 7.0).all() and (p16[:, spec.num_priv:] == 7.0).all(), "padding columns untouched"


def test_two_phase_equals_fused():
    """REWARD and OBS phases launched separately (as a real Isaac Gym env does around its gym reset calls)
    give the same result as the single fused launch."""
    N = 512
    cfg = XBotMathConfig.for_dof(12, num_envs=N)
    g = torch.Generator(device=DEV).manual_seed(5)
    a, b = XBotEnvMath(cfg, DEV), XBotEnvMath(cfg, DEV)
    ep = torch.randint(0, 2400, (N,), device=DEV, generator=g)
    a.episode_length_buf.copy_(ep)
    b.episode_length_buf.copy_(ep)
    for s in range(18):
        sim = synthetic_sim_state(cfg, g, DEV, upright=True, base_contact_frac=0.05)
        sim["actions"] = torch.randn(N, 12, device=DEV, generator=g)
        noise = torch.randn(N, cfg.num_single_obs, device=DEV, generator=g)
        oa = [t.clone() for t in a.step(sim, noise)]
        b.step(sim, noise, phases=ENV_PHASE_REWARD)
        ob = b.step(sim, noise, phases=ENV_PHASE_OBS)
        for x, y in zip(oa, ob):
            assert torch.equal(x, y), f"step {s}"
    assert torch.equal(a.episode_sums_tensor, b.episode_sums_tensor)


def test_in_kernel_noise_statistics():
    """noise=None -> Philox4x32-10 in the kernel: mean ~0, std ~ noise_scale_vec * noise_level, fresh per step."""
    N = 8192
    cfg = XBotMathConfig.for_dof(12, num_envs=N)
    g = torch.Generator(device=DEV).manual_seed(9)
    noisy, clean = XBotEnvMath(cfg, DEV), XBotEnvMath(XBotMathConfig(num_envs=N, add_noise=False), DEV)
    sim = synthetic_sim_state(cfg, g, DEV)
    sim["actions"] = torch.zeros(N, 12, device=DEV)
    o1 = noisy.step(sim, None)[0][:, -47:].clone()
    o2 = noisy.step(sim, None)[0][:, -47:].clone()
    c = clean.step(sim, None)[0][:, -47:]
    d1, d2 = (o1 - c), (o2 - c)
    sigma = (cfg.noise_scale_vec() * cfg.noise_level).to(DEV)
    on = sigma > 0
    assert torch.all(d1[:, ~on] == 0)
    assert torch.allclose(d1[:, on].std(0), sigma[on], rtol=0.05)
    assert float(d1[:, on].mean(0).abs().max()) < 4 * float(sigma.max()) / N ** 0.5
    assert not torch.equal(d1, d2)


# ---------------------------------------------------------------------------------- SURVEY 8f rank 2
CB_SIM_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_state", "torques", "env_frictions", "body_mass")


def test_callback_golden_replay(golden_dir):
    """`_post_physics_step_callback` (+ `_resample_commands`, `_push_robots`) and the command / extras part of `reset_idx`
    inside the stateful env loop, against the fixture recorded from the UNMODIFIED reference with replayed uniform draws.
    Exact: flags, which envs were resampled / pushed, the pushed root velocities and push buffers (pure fp32 mul + add).
    <= 1e-5 relative: the heading-derived yaw command (atan2f vs libm), observations, rewards, episode means."""
    z = np.load(f"{golden_dir}/env_callback_d12.npz")
    N, steps = int(z["N"]), int(z["steps"])
    env = XBotEnvMath(XBotMathConfig.for_dof(12, num_envs=N), DEV)
    env.episode_length_buf.copy_(torch.tensor(z["ep_len0"]))
    env._callback_state()
    env.commands.copy_(torch.tensor(z["commands0"]))
    env.common_step_counter = int(z["counter0"])
    t = lambda k, s: torch.tensor(z[k][s]).to(DEV).contiguous()
    pushes = means = 0
    for s in range(steps):
        sim = {k: t(k, s) for k in CB_SIM_KEYS}
        sim["actions"] = t("actions", s)
        obs, priv, rew, reset, tout, episode = env.step_with_callbacks(sim, t("noise", s), u_cmd=t("u_cmd", s), u_push=t("u_push", s),
                                                                       u_cmd_reset=t("u_cmd_reset", s))
        torch.cuda.synchronize()
        assert torch.equal(reset.cpu(), torch.tensor(z["reset"][s])) and torch.equal(tout.cpu(), torch.tensor(z["time_out"][s])), s
        want_cmd = torch.tensor(z["commands"][s])
        got_cmd = env.commands.cpu()
        assert torch.equal(got_cmd[:, [0, 1, 3]], want_cmd[:, [0, 1, 3]]), f"step {s}: resampled commands"
        assert close(got_cmd[:, 2], want_cmd[:, 2])[0], f"step {s}: heading -> yaw command {close(got_cmd[:, 2], want_cmd[:, 2])[1]}"
        assert torch.equal(sim["root_states"].cpu(), torch.tensor(z["root_states_out"][s])), f"step {s}: pushed root states"
        assert torch.equal(env.rand_push_force.cpu(), torch.tensor(z["rand_push_force"][s]))
        assert torch.equal(env.rand_push_torque.cpu(), torch.tensor(z["rand_push_torque"][s]))
        for name, got, want in (("obs", obs, z["obs"][s]), ("priv", priv, z["priv"][s]), ("rew", rew, z["rew"][s])):
            ok, err = close(got.cpu(), torch.tensor(want))
            assert ok, f"step {s} {name}: max abs err {err}"
        pushes += int(z["pushed"][s])
        assert int(env.episode_reset_count.item()) == int(z["reset"][s].sum())
        if z["has_means"][s]:
            means += 1
            ok, err = close(env.episode_means_tensor.cpu(), torch.tensor(z["episode_means"][s]))
            assert ok, f"step {s}: episode means {err}"
            assert set(episode) == {"rew_" + k for k in eo.xbot_spec(12).active_rewards()}
    assert pushes >= 1 and means >= 3
    ok, err = close(env.episode_sums_tensor.cpu(), torch.tensor(z["final_episode_sums"]))
    assert ok, err


def test_callback_vs_oracle_4096_and_philox():
    """4096 envs x 6 steps with a push on step 2 against the oracle (given draws), then the in-kernel Philox draws:
    resampled commands stay inside their ranges, are zeroed below 0.2 and untouched elsewhere."""
    N = 4096
    spec, cb = eo.xbot_spec(12), eo.CallbackSpec()
    cfg = XBotMathConfig.for_dof(12, num_envs=N)
    env = XBotEnvMath(cfg, DEV)
    st = eo.EnvState(spec, N)
    g = torch.Generator().manual_seed(11)
    ep0 = torch.randint(0, 2300, (N,), generator=g)
    ep0[:300] = 798                                            # a block of envs hits the resampling boundary on step 2
    st.episode_length_buf = ep0.clone()
    env.episode_length_buf.copy_(ep0)
    env._callback_state()
    cmd0 = torch.cat((0.5 * torch.randn(N, 3, generator=g), 3.0 * (2 * torch.rand(N, 1, generator=g) - 1)), dim=1)
    st.commands, st.rand_push_force, st.rand_push_torque = cmd0.clone(), torch.zeros(N, 3), torch.zeros(N, 3)
    env.commands.copy_(cmd0)
    env.common_step_counter = 398
    for s in range(6):
        sim = eo.synthetic_sim_state(spec, N, g, base_contact_frac=0.03)
        noise = torch.randn(N, spec.So, generator=g)
        u_cmd, u_push, u_rst = torch.rand(N, 3, generator=g), torch.rand(N, 5, generator=g), torch.rand(N, 3, generator=g)
        want = eo.env_step(spec, st, sim, noise, callback=dict(spec=cb, common_step_counter=398 + s + 1, u_cmd=u_cmd, u_push=u_push,
                                                               u_cmd_reset=u_rst))
        dsim = {k: sim[k].to(DEV).contiguous() for k in CB_SIM_KEYS}
        dsim["actions"] = sim["actions"].to(DEV)
        obs, priv, rew, reset, tout, _ = env.step_with_callbacks(dsim, noise.to(DEV), u_cmd=u_cmd.to(DEV), u_push=u_push.to(DEV),
                                                                 u_cmd_reset=u_rst.to(DEV))
        assert want["pushed"] == (s == 1)
        assert torch.equal(reset.cpu(), want["reset"])
        assert torch.equal(env.commands.cpu()[:, [0, 1, 3]], want["commands"][:, [0, 1, 3]])
        assert close(env.commands.cpu()[:, 2], want["commands"][:, 2])[0]
        assert torch.equal(dsim["root_states"].cpu(), want["root_states"])
        for name, a, b in (("obs", obs, want["obs"]), ("priv", priv, want["priv"]), ("rew", rew, want["rew"])):
            ok, err = close(a.cpu(), b)
            assert ok, (s, name, err)
        if want["episode_means"] is not None:
            ok, err = close(env.episode_means_tensor.cpu(), want["episode_means"])
            assert ok, (s, err)
    # in-kernel Philox draws
    env.episode_length_buf.fill_(799)
    before = env.commands.clone()
    env.episode_length_buf[N // 2:] = 5
    root = dsim["root_states"].clone()
    env.common_step_counter = 799                              # -> 800: a push step
    assert env.post_physics_callback(root) is True
    c = env.commands.cpu()
    half = N // 2
    assert torch.equal(c[half:, [0, 1, 3]], before.cpu()[half:, [0, 1, 3]])          # not at the boundary: untouched
    x, y, h = c[:half, 0], c[:half, 1], c[:half, 3]
    nz = (x != 0) | (y != 0)
    assert 0.3 < nz.float().mean() < 1.0
    assert (x[nz] >= -0.3).all() and (x[nz] <= 0.6).all() and (y[nz] >= -0.3).all() and (y[nz] <= 0.3).all()
    assert (torch.sqrt(x[nz] ** 2 + y[nz] ** 2) > 0.2).all()
    assert (h >= -3.14).all() and (h <= 3.14).all() and h.std() > 1.0
    assert (c[:, 2].abs() <= 1.0).all()
    pf, pt = env.rand_push_force.cpu(), env.rand_push_torque.cpu()
    assert (pf[:, :2].abs() <= 0.2).all() and (pt.abs() <= 0.4).all() and pf[:, :2].std() > 0.05 and pt.std() > 0.1
    assert torch.equal(root.cpu()[:, 7:9], pf[:, :2]) and torch.equal(root.cpu()[:, 10:13], pt)
    assert torch.equal(env.root_vel_prepush.cpu(), dsim["root_states"].cpu()[:, 7:13])


def test_synthetic_env_with_callbacks_is_a_complete_vecenv():
    """`SyntheticXBotEnv(callbacks=True)`: the VecEnv loop with command resampling, pushes and extras["episode"]."""
    from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv
    torch.manual_seed(0)
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=512), DEV, callbacks=True)
    env.math._callback_state()
    env.math.common_step_counter = 397
    acts = torch.zeros(512, 12, device=DEV)
    seen_push = False
    for i in range(5):
        obs, priv, rew, dones, extras = env.step(acts)
        seen_push = seen_push or bool(env.math.rand_push_force.abs().sum() > 0)
        assert obs.shape == (512, 705) and priv.shape == (512, 219) and rew.shape == (512,)
        assert set(extras) >= {"time_outs", "episode"} and len(extras["episode"]) == 22
        assert all(torch.isfinite(v) for v in extras["episode"].values())
    assert seen_push and torch.isfinite(obs).all() and (env.math.commands[:, 2].abs() <= 1).all()


# ---------------------------------------------------------------------------------- fp32 phase precision
def _one_step(cfg, ep_after, D=12):
    """One fused step at N = len(ep_after) envs whose episode length AFTER the increment of :780 is `ep_after`; no
    contacts on the base (no termination).  Returns (obs, priv) on the CPU and the sim dict."""
    N = len(ep_after)
    env = XBotEnvMath(cfg, DEV)
    env.episode_length_buf.copy_(torch.as_tensor(ep_after, dtype=torch.long) - 1)
    g = torch.Generator(device=DEV).manual_seed(3)
    sim = synthetic_sim_state(cfg, g, DEV, base_contact_frac=0.0)
    sim["actions"] = torch.randn(N, D, device=DEV, generator=g)
    obs, priv, rew, reset, tout = env.step(sim, None if not cfg.add_noise else torch.zeros(N, cfg.num_single_obs, device=DEV))
    assert not bool(reset.any())
    return obs.cpu(), priv.cpu(), {k: v.cpu() for k, v in sim.items()}


def test_phase_mul_reciprocal_reproduces_torch_cuda_division():
    """`phase = episode_length_buf * dt / cycle_time` (:692): torch-CPU divides, torch-CUDA multiplies by the fp32
    reciprocal of the Python scalar -- a <= 1-ulp difference in the phase that sin/cos amplify to ~2e-5 at phase ~37.
    `phase_mul_reciprocal=False` must follow the CPU semantics (the oracle's), `True` the CUDA ones: checked against
    torch's own CUDA kernels evaluating the reference's expression on this GPU, for EVERY episode length 1..2399."""
    ep = torch.arange(1, 2400)
    outs = {}
    for mode in (False, True):
        cfg = XBotMathConfig(num_envs=len(ep), add_noise=False, phase_mul_reciprocal=mode)
        obs, _, _ = _one_step(cfg, ep)
        outs[mode] = obs[:, -47:-45].clone()                      # newest frame: sin, cos
    cpu_phase = ep * 0.01 / 0.64                                  # torch-CPU: int64 * float -> fp32, true division
    cuda_phase = ep.to(DEV) * 0.01 / 0.64                         # torch-CUDA: the reference's expression on the device
    want = {False: torch.stack((torch.sin(2 * torch.pi * cpu_phase), torch.cos(2 * torch.pi * cpu_phase)), 1),
            True: torch.stack((torch.sin(2 * torch.pi * cuda_phase), torch.cos(2 * torch.pi * cuda_phase)), 1).cpu()}
    n_phase_differs = int((cpu_phase != cuda_phase.cpu()).sum())
    print(f"episode lengths whose fp32 phase differs between torch-CPU and torch-CUDA: {n_phase_differs} of {len(ep)}")
    for mode in (False, True):
        err = float((outs[mode] - want[mode]).abs().max())
        print(f"phase_mul_reciprocal={mode}: max |sin/cos - torch-{'CUDA' if mode else 'CPU'}| = {err:.3e}")
        assert err <= 3e-7, (mode, err)                           # sinf / cosf implementations: <= 2 ulp
    cross = float((outs[True] - want[False]).abs().max())
    print(f"cross-check (CUDA semantics vs CPU reference): {cross:.3e}")
    if n_phase_differs:                                           # the two semantics are really different computations
        assert cross > 1e-6 and not torch.equal(outs[False], outs[True])


def test_dead_band_decisions_at_adversarial_phases():
    """The |sin| < 0.1 dead band (:713, :742) and the sin >= 0 / < 0 split (:707-711) flip whole rows of `ref_dof_pos` and
    the stance mask.  For cycle times chosen so that an integer episode length lands on the fp32 phases NEAREST to a
    band edge (both sides, within a few ulp of the phase), the kernel's decisions must (a) equal the oracle's and (b) be
    exactly the function of the sin value the kernel itself reports: stance, ref_dof_pos (through priv's q - ref columns)."""
    import math
    cases = []
    for ep_t, k in ((7, 0), (37, 1), (411, 12), (1999, 61)):      # edge theta = k*pi + asin(0.1): early, mid and late in an episode
        t32 = np.float32(ep_t) * np.float32(0.01)
        theta = k * math.pi + (math.asin(0.1) if k % 2 == 0 else -math.asin(0.1))
        c0 = np.float32(2 * math.pi * float(t32) / theta)
        cs = [c0]
        up, dn = c0, c0
        for _ in range(256):
            up, dn = np.nextafter(up, np.float32(10)), np.nextafter(dn, np.float32(0))
            cs += [up, dn]
        cs = np.array(cs, dtype=np.float32)
        s = torch.sin(torch.tensor(np.float32(6.2831855) * (t32 / cs)))          # the oracle's arithmetic (:692, :705)
        margin = (s.abs() - 0.1).abs()
        order = torch.argsort(margin)[:24]                         # the 24 cycle times whose sin is closest to the edge
        assert float(margin[order[0]]) < 2e-5
        cases += [(ep_t, float(cs[i])) for i in order.tolist()]
    below = above = exempt = 0
    for ep_t, c in cases:
        spec = eo.xbot_spec(12)
        spec.cycle_time, spec.add_noise = c, False
        cfg = XBotMathConfig(num_envs=4, add_noise=False, cycle_time=c)
        eps = [ep_t, ep_t, ep_t + 32, ep_t + 1]
        obs, priv, sim = _one_step(cfg, eps)
        st = eo.EnvState(spec, 4)
        st.episode_length_buf = torch.tensor(eps) - 1
        sim_o = dict(sim)
        D = 12
        sim_o["dof_pos"], sim_o["dof_vel"] = sim["dof_state"].view(4, D, 2)[..., 0], sim["dof_state"].view(4, D, 2)[..., 1]
        want = eo.env_step(spec, st, sim_o, None)
        pf, wf = priv[:, -73:], want["priv"][:, -73:]              # newest critic frame
        s_gpu = pf[:, 0]
        assert torch.allclose(s_gpu, wf[:, 0], atol=3e-7), (ep_t, c)
        # (a) same decisions as the oracle -- except where the CPU's own sin is within 2 ulp of the edge value: there
        # CUDA sinf and glibc sinf (each <= 2 ulp) may legitimately land on different sides
        far = ((wf[:, 0].abs() - 0.1).abs() > 1.6e-8)
        exempt += int((~far).sum())
        assert torch.equal(pf[far, 69:71], wf[far, 69:71]), f"stance mask differs from the oracle at ep {ep_t}, cycle_time {c!r}"
        assert torch.allclose(pf[far, 41:53], wf[far, 41:53], atol=1e-6), f"q - ref differs at ep {ep_t}, cycle_time {c!r}"
        # (b) decisions are exactly the function of the kernel's own sin value
        dead = s_gpu.abs() < 0.1
        stance = torch.stack(((s_gpu >= 0) | dead, (s_gpu < 0) | dead), 1).float()
        assert torch.equal(pf[:, 69:71], stance)
        q = sim["dof_state"].view(4, D, 2)[..., 0]
        ref = q - pf[:, 41:53]
        sl, sr = torch.clamp(s_gpu, max=0.0), torch.clamp(s_gpu, min=0.0)
        exp = torch.zeros(4, D)
        exp[:, 2], exp[:, 3], exp[:, 4] = sl * 0.17, sl * 0.34, sl * 0.17
        exp[:, 8], exp[:, 9], exp[:, 10] = sr * 0.17, sr * 0.34, sr * 0.17
        exp[dead] = 0.0
        assert torch.allclose(ref, exp, atol=2e-7), (ep_t, c)
        assert torch.equal(ref[dead], torch.zeros_like(ref[dead]) + 0.0) or float(ref[dead].abs().max()) < 2e-7
        below += int(dead[0])
        above += int(~dead[0])
    print(f"{len(cases)} adversarial (episode length, cycle time) cases: {below} inside the dead band, {above} outside, "
          f"{exempt} env rows exactly on the edge (oracle comparison skipped, self-consistency still checked)")
    assert below >= 8 and above >= 8


def test_compact_body_tables_are_bit_identical_to_full_tables():
    """SyntheticXBotEnv(compact_bodies=True) ships only the rigid-body rows (2 feet + 2 knees) and contact rows (base + feet)
    the kernel reads; with the same pool seed every output must equal the full-table run bit for bit, resets included."""
    N = 200
    envs = [SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), DEV, pool=3, seed=7, torch_rng_noise=False, compact_bodies=c)
            for c in (False, True)]
    assert envs[1].pool[0]["rigid_state"].shape == (N, 4, 13) and envs[1].pool[0]["contact_forces"].shape[1] == 3
    g = torch.Generator(device=DEV).manual_seed(3)
    resets = 0
    for s in range(12):
        a = torch.randn(N, 12, device=DEV, generator=g)
        outs = [e.step(a.clone()) for e in envs]
        for x, y, name in zip(outs[0][:4], outs[1][:4], ("obs", "priv", "rew", "reset")):
            assert torch.equal(x, y), (s, name)
        assert torch.equal(outs[0][4]["time_outs"], outs[1][4]["time_outs"])
        resets += int(outs[0][3].sum())
    assert resets > 0
    for name in ("feet_air_time", "feet_height", "last_feet_z", "episode_sums_tensor", "last_contacts"):
        assert torch.equal(getattr(envs[0].math, name), getattr(envs[1].math, name)), name
