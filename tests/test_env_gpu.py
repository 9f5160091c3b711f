"""Fused env-step kernel vs the golden vectors produced by the UNMODIFIED reference env math and vs the oracle (-m gpu).

Tolerances: reset / time_out flags and history indexing (which frame lands where, what is zeroed) are exact;
fp32 observations / rewards agree within 1e-5 relative (+1e-6 absolute for values near zero: CUDA's sinf/cosf/expf/
atan2f differ from the CPU libm by <= 2 ulp)."""
import numpy as np
import pytest
import torch

from gpu_util import DEV
from humanoid_gym_with_comments_b200._lib import ENV_PHASE_OBS, ENV_PHASE_REWARD
from humanoid_gym_with_comments_b200.envs import XBotEnvMath, XBotMathConfig, synthetic_sim_state
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
