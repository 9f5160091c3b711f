"""History ring -- the env writes ONE new frame per step, stacked observations are windows of an
env-major ring -- against the golden vectors of the unmodified reference, against the stacked format (bit-exact:
history-buffer indexing is integer work), and through the learner against the oracle (-m gpu).

Reference semantics reproduced: obs_history / critic_history deques re-stacked every step (humanoid_env.py:880-887), all
history frames of a reset env zeroed BEFORE the new frame is appended (reset_idx :1159-1163), +-18 clip (:654-657)."""
import numpy as np
import pytest
import torch

from gpu_util import DEV
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic, RolloutStorage
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotEnvMath, XBotMathConfig, synthetic_sim_state
from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu
SIM_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_state", "commands", "torques", "rand_push_force",
            "rand_push_torque", "env_frictions", "body_mass")
NETS = dict(actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])


def ring_storage(cfg, T):
    st = RolloutStorage(cfg.num_envs, T, [cfg.num_observations], [cfg.num_privileged_obs], [cfg.num_dof], DEV)
    st.enable_ring(cfg.num_single_obs, cfg.frame_stack, cfg.single_num_privileged_obs, cfg.c_frame_stack)
    return st


def check_bf16_window(ring, row, stacked):
    """live16 window of `row` == bf16(stacked fp32 row) at the frame pitch, pad columns zero."""
    w16 = ring.view16(row).float().view(-1, ring.frames, ring.pitch16)
    want = stacked.view(-1, ring.frames, ring.width).to(torch.bfloat16).float()
    assert torch.equal(w16[:, :, :ring.width], want)
    assert not bool(w16[:, :, ring.width:].any())


@pytest.mark.parametrize("D", [12, 18])
def test_ring_golden_replay(D, golden_dir):
    """The reference's own 20 stateful steps (24 resets, 2 time-outs) through the ring format, T = 7 rows per iteration so
    the window is handed over to the next iteration twice.  Live views vs the golden obs / priv (same tolerance as the
    stacked replay); the rows re-stacked from the archive for the update must equal the live views bit for bit even after
    LATER resets zeroed the live history."""
    z = np.load(f"{golden_dir}/env_d{D}.npz")
    N, steps, T = int(z["N"]), int(z["steps"]), 7
    cfg = XBotMathConfig.for_dof(D, num_envs=N)
    env = XBotEnvMath(cfg, DEV)
    env.episode_length_buf.copy_(torch.tensor(z["ep_len0"]))
    st = ring_storage(cfg, T)
    live = {}                                    # row -> (obs, priv) clones of the live views, this iteration
    for s in range(steps):
        sim = {k: torch.tensor(z[k][s]).to(DEV).contiguous() for k in SIM_KEYS}
        sim["actions"] = torch.tensor(z["actions"][s]).to(DEV)
        row = st.step + 1
        obs, priv, rew, reset, tout = env.step(sim, torch.tensor(z["noise"][s]).to(DEV), ring=(st, row))
        assert obs.shape == (N, cfg.num_observations) and obs.data_ptr() == st._ring_obs.ptr32(row)
        assert torch.equal(reset.cpu(), torch.tensor(z["reset"][s])) and torch.equal(tout.cpu(), torch.tensor(z["time_out"][s]))
        for name, got, want in (("obs", obs, z["obs"][s]), ("priv", priv, z["priv"][s]), ("rew", rew, z["rew"][s])):
            assert torch.allclose(got.cpu(), torch.tensor(want), rtol=1e-5, atol=1e-6), (D, s, name)
        zero_ref = torch.tensor(z["obs"][s]) == 0      # history indexing is exact: every reset-zeroed frame is zero here too
        assert torch.equal((obs.cpu() == 0) | ~zero_ref, torch.ones_like(zero_ref)), f"step {s}: zeroed history frames"
        check_bf16_window(st._ring_obs, row, obs)
        check_bf16_window(st._ring_priv, row, priv)
        live[row] = (obs.clone(), priv.clone())
        st.step += 1
        if st.step == T or s == steps - 1:
            mo, mp = st.observations, st.privileged_observations           # re-stacked from the archive ring
            for r, (o, p) in live.items():
                if r < T:
                    assert torch.equal(mo[r], o) and torch.equal(mp[r], p), f"archive row {r} (global step {s})"
            last = live[st.step]
            st.clear()                                                    # window of the last row -> row 0
            assert torch.equal(st._ring_obs.view32(0), last[0]) and torch.equal(st._ring_priv.view32(0), last[1])
            check_bf16_window(st._ring_obs, 0, last[0])
            assert torch.equal(st.observations[0], last[0])
            live = {0: last}


@pytest.mark.parametrize("N,T,ITERS,min_resets", [(4096, 12, 3, 500), (5003, 6, 2, 200)])
def test_ring_equals_stacked(N, T, ITERS, min_resets):
    """3 iterations x 12 steps with ~3 % resets per step: the ring's live views and the rows re-stacked from its archive are
    BIT-IDENTICAL to the stacked format (same kernel math, different history plumbing), as are rewards, flags and every
    piece of persistent env state.  4096 envs run the 16-envs-per-block ring kernel, 5003 the 32-envs-per-block one (more
    than two blocks per SM; ~60 KB of staged inputs per block, a ragged last block)."""
    cfg = XBotMathConfig.for_dof(12, num_envs=N)
    g = torch.Generator(device=DEV).manual_seed(21)
    a, b = XBotEnvMath(cfg, DEV), XBotEnvMath(cfg, DEV)
    ep = torch.randint(0, 2400, (N,), device=DEV, generator=g)
    a.episode_length_buf.copy_(ep)
    b.episode_length_buf.copy_(ep)
    st = ring_storage(cfg, T)
    resets = 0
    for it in range(ITERS):
        stacked = []
        for t in range(T):
            sim = synthetic_sim_state(cfg, g, DEV, upright=(t % 2 == 0), base_contact_frac=0.03)
            sim["actions"] = torch.randn(N, 12, device=DEV, generator=g)
            noise = torch.randn(N, cfg.num_single_obs, device=DEV, generator=g)
            oa, pa, ra, da, ta = [x.clone() for x in a.step(sim, noise)]
            ob, pb, rb, db, tb = b.step(sim, noise, ring=(st, st.step + 1))
            assert torch.equal(oa, ob) and torch.equal(pa, pb), (it, t)
            assert torch.equal(ra, rb) and torch.equal(da, db) and torch.equal(ta, tb)
            resets += int(da.sum())
            stacked.append((oa, pa))
            st.step += 1
        mo, mp = st.observations, st.privileged_observations
        for r in range(1, T):
            assert torch.equal(mo[r], stacked[r - 1][0]) and torch.equal(mp[r], stacked[r - 1][1]), (it, r)
        st.clear()
    for name in ("actions", "last_actions", "last_last_actions", "last_dof_vel", "last_root_vel", "feet_air_time", "feet_height",
                 "last_feet_z", "ref_dof_pos", "episode_length_buf", "episode_sums_tensor"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    assert resets > min_resets


def make_alg(N, T, seed=0):
    torch.manual_seed(seed)
    ac = ActorCritic(705, 219, 12, init_noise_std=1.0, **NETS)
    c = po.xbot_ppo_config()
    alg = PPO(ac, num_learning_epochs=c.num_learning_epochs, num_mini_batches=c.num_mini_batches, clip_param=c.clip_param,
              gamma=c.gamma, lam=c.lam, value_loss_coef=c.value_loss_coef, entropy_coef=c.entropy_coef,
              learning_rate=c.learning_rate, max_grad_norm=c.max_grad_norm, use_clipped_value_loss=True,
              schedule=c.schedule, desired_kl=c.desired_kl, device=DEV)
    alg.init_storage(N, T, [705], [219], [12])
    return alg, c


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def test_ring_iteration_vs_oracle():
    """BASELINE config 1 shapes (64 envs x 24 steps) through the whole ring path: env frames -> ring -> layer 1 reads the
    bf16 windows in place (frame-pitched first-layer weights) -> update gathers the minibatch rows out of the archive ring
    and maps the frame-pitched weight gradient back -> every minibatch against the oracle fed with the SAME rollout
    (observations re-stacked from the ring).  Two iterations, so the hand-over of the window is on the path as well."""
    N, T = 64, 24
    alg, cfg = make_alg(N, T)
    torch.manual_seed(0)
    p = po.init_params(705, 219, 12, NETS["actor_hidden_dims"], NETS["critic_hidden_dims"], (128, 128), 1.0)
    adam, lr = po.AdamState(), cfg.learning_rate
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), DEV, seed=4)
    env.bind_storage(alg.storage, gamma=cfg.gamma)
    assert alg.storage.ring_enabled
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    gen = torch.Generator().manual_seed(8)
    for it in range(2):
        with torch.inference_mode():
            for t in range(T):
                a = alg.act(obs, cobs)
                obs, cobs, rew, dones, infos = env.step(a)
                alg.process_env_step(rew, dones, infos)
            alg.compute_returns(cobs)
        st = alg.storage
        cpu = lambda x: x.detach().cpu().clone()
        ro = po.Rollout(cpu(st.observations), cpu(st.privileged_observations), cpu(st.actions), cpu(st.rewards), cpu(st.dones),
                        cpu(st.values), cpu(st.actions_log_prob), cpu(st.mu), cpu(st.sigma))
        ro.returns, ro.advantages = cpu(st.returns), cpu(st.advantages)
        assert float(ro.observations.abs().max()) <= 18.0 and float(ro.observations[1:].abs().sum()) > 0
        perm = torch.randperm(N * T, generator=gen)
        alg.perm_override = perm.to(DEV)
        rec = []
        ov, os_, olv, lr = po.ppo_update(p, adam, ro, perm, cfg, lr, record=rec)
        v, s, _, lv = alg.update()
        log = alg.last_update_log.numpy()
        for i, r in enumerate(rec):
            print(f"it{it} mb{i}: value {log[i,0]:.5f}/{r['value']:.5f} surr {log[i,1]:+.5e}/{r['surrogate']:+.5e} "
                  f"lv {log[i,2]:.5f}/{r['lin_vel']:.5f} kl {log[i,3]:.3e}/{r['kl']:.3e} gn {log[i,6]:.4f}/{r['grad_norm']:.4f}")
            tol = 2e-3 if it == 0 else 5e-3
            assert rel(log[i, 0], r["value"]) <= tol and rel(log[i, 2], r["lin_vel"]) <= tol, (it, i)
            assert abs(log[i, 1] - r["surrogate"]) <= 2e-3
            assert rel(log[i, 4], r["entropy"]) <= 1e-5
            assert log[i, 5] == pytest.approx(r["lr"], rel=1e-12)
            assert rel(log[i, 6], r["grad_norm"]) <= 2e-2
    sd = alg.actor_critic.state_dict()
    torch.manual_seed(0)
    w0 = po.init_params(705, 219, 12, NETS["actor_hidden_dims"], NETS["critic_hidden_dims"], (128, 128), 1.0)["actor.0.weight"]
    cos = torch.nn.functional.cosine_similarity((sd["actor.0.weight"].cpu() - w0).flatten(), (p["actor.0.weight"] - w0).flatten(), dim=0)
    print("cosine(update direction) actor.0.weight:", float(cos))
    assert float(cos) > 0.99
    # the frame-pitched weight shadow is kept in step with the fp32 master by the optimiser launch
    flat = alg.actor_critic.flat
    w = alg.actor_critic.actor[0].weight
    sh = flat.w16_framed[(id(w), 47, 48)]
    want = torch.zeros_like(sh)
    want.view(512, -1)[:, :720].view(512, 15, 48)[:, :, :47] = w.data.view(512, 15, 47).to(torch.bfloat16)
    assert torch.equal(sh, want)


def test_ring_and_stacked_training_loops_agree():
    """The same bound env + learner loop in both formats (identical seeds, in-kernel Philox): layer 1 sums its 705 products in
    a different order (frame pitch 48 vs 47), so outputs agree to bf16 rounding noise, not bit for bit."""
    N, T = 512, 8
    outs = {}
    for ring in (True, False):
        torch.manual_seed(3)
        env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), DEV, seed=1, torch_rng_noise=False)
        alg, cfg = make_alg(N, T)
        alg.torch_rng = False
        env.bind_storage(alg.storage, gamma=cfg.gamma, ring=ring)
        assert alg.storage.ring_enabled == ring
        obs, cobs = env.get_observations(), env.get_privileged_observations()
        res = []
        for it in range(2):
            with torch.inference_mode():
                for t in range(T):
                    a = alg.act(obs, cobs)
                    obs, cobs, rew, dones, infos = env.step(a)
                    alg.process_env_step(rew, dones, infos)
                alg.compute_returns(cobs)
            vals = alg.storage.values.clone()
            res.append((vals, alg.storage.observations.clone(), alg.update()))
        outs[ring] = res
    for it in range(2):
        (va, oa, la), (vb, ob, lb) = outs[True][it], outs[False][it]
        assert torch.allclose(va, vb, atol=5e-3), float((va - vb).abs().max())
        assert torch.allclose(oa, ob, atol=5e-3)                 # observations carry the (slightly different) actions
        assert rel(la[0], lb[0]) <= 5e-3 and rel(la[3], lb[3]) <= 5e-3 and abs(la[1] - lb[1]) <= 2e-3, (la, lb)


def test_ring_rejects_foreign_observations_and_unbinds():
    from humanoid_gym_with_comments_b200._lib import HgError
    N, T = 64, 4
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), DEV)
    alg, cfg = make_alg(N, T)
    env.bind_storage(alg.storage, gamma=cfg.gamma)
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    alg.act(obs, cobs)                                          # the ring's own window: fine
    with pytest.raises(HgError, match="history ring"):
        alg.act(obs.clone(), cobs.clone())                      # a copy is not the ring
    alg.storage.clear()
    env.bind_storage(None)                                      # back to the env's own stacked buffers, history preserved
    assert not alg.storage.ring_enabled
    o2 = env.get_observations()
    assert torch.equal(o2, obs)
    a = alg.act(o2.clone(), env.get_privileged_observations().clone())   # generic path works again
    assert a.shape == (N, 12)
