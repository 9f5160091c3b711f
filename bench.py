#!/usr/bin/env python
"""bench.py -- PPO env-steps/s (rollout + update, device-timed) of the B200-native hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs 4096] [--horizon 60]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one PPO iteration of BASELINE.json configs[1] (XBot-L humanoid_ppo defaults: 4096 envs x 60 steps per GPU,
actor [512,256,128] / critic [768,256,128] / lin-vel [128,128], 2 epochs x 4 minibatches): 60 x (fused env step + act +
process_env_step) + compute_returns + update, driven through the reference-facing PPO / VecEnv API exactly as
OnPolicyRunner.learn does (on_policy_runner.py:125-168).  Isaac Gym's PhysX step is replaced by synthetic
Isaac-Gym-shaped state tensors.  One JSON line is printed by rank 0.

  value  : whole-job env-steps/s with the synthetic state already resident in HBM
  e2e    : same loop, but the Isaac-Gym-shaped state of every env step arrives from pinned host memory (12 steps per copy on
           a side stream, inside the timed region) and every
           iteration reads its losses + mean reward back to the host
  roofline: the grouped tcgen05 GEMM launches (the dominant kernel), algorithmic FLOPs / CUDA-event time, against the
           measured bf16 peak of MEASURED_PEAKS.json
  cpu_baseline: the UNMODIFIED reference (oracle/_ref; the oracle port only if that is absent) on torch-CPU fp32, all host
           threads, on a bounded sample of the same workload
  reference_torch_cuda: the same unmodified reference on cuda:0 at the FULL workload (eager torch fp32, CUDA events)
  --impl reference: the reference arm on the host cores; its config names the envs it actually ran
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ppo_env_steps_per_sec"
UNIT = "env-steps/s"
NETS = dict(actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
FALLBACK_BF16_TFLOPS = 1400.0     # B200_PROFILING.md fallback (sustained)
FALLBACK_HBM_GBS = 6650.0       # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU")
    ap.add_argument("--horizon", type=int, default=60)
    ap.add_argument("--cpu-sample-envs", type=int, default=256, help="envs of the cpu_baseline sample inside the product line")
    ap.add_argument("--ref-envs", type=int, default=0, help="--impl reference: envs per iteration (0 = the full workload if "
                    "steps + warmup iterations fit ~150 s on this host, else the largest power-of-two sample that does)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip timing the unmodified reference on cuda:0 (reference_torch_cuda)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--torch-rng", action="store_true",
                    help="draw action / observation noise with torch.randn (stream parity with the reference on the same device) instead of the in-kernel Philox4x32-10")
    ap.add_argument("--unbound-env", action="store_true", help="env writes its own buffers; PPO.act ingests them (generic drop-in path)")
    ap.add_argument("--full-body-tables", action="store_true",
                    help="synthetic simulator state with all 13 rigid bodies / contact rows (1,076 B per env-step over PCIe in the "
                         "e2e arm) instead of only the rows the kernel reads (2 feet + 2 knees; base + feet: 488 B)")
    ap.add_argument("--stacked", action="store_true", help="bound env writes whole stacked rows (round-1 format) instead of the history ring")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: 'cuda' runs the same torch port on cuda:0 at the FULL workload (the torch-CUDA "
                         "figure north_star's target is quoted against); the default and the driver's arm is 'cpu'")
    return ap.parse_args()


def workload_config(a, world):
    return {"workload": f"XBot-L humanoid_ppo: {a.envs} envs/GPU x {a.horizon} steps, obs 705 / priv 219 / 12 actions, "
                        f"actor [512,256,128] critic [768,256,128] lin-vel [128,128], 2 epochs x 4 minibatches "
                        f"(BASELINE.json configs[1]); synthetic Isaac-Gym-shaped state, fused env step in the loop",
            "envs_per_gpu": a.envs, "horizon": a.horizon, "parallelism": f"env-sharded dp{world}",
            "rng": "in-kernel Philox4x32-10 keyed by (seed, step, row) for action and observation noise; --torch-rng switches to torch.randn",
            "precision": "bf16 tensor-core GEMM operands with fp32 accumulation (hidden layers); output layers, losses, GAE, env math and Adam in fp32",
            "l2_policy": "per-iteration working set (history rings 0.34 GB, permuted bf16 minibatch rows 0.46 GB, activations / "
                         "gradients ~0.6 GB, all rewritten every iteration) exceeds the 126 MB L2; no explicit flush",
            "obs_format": "history ring (one new 47 / 73-wide frame per env step; SURVEY 8f rank 3)",
            "sim_state": ("full Isaac Gym tables (13 rigid bodies, 13 contact rows): 1,076 B per env-step" if getattr(a, "full_body_tables", False)
                          else "only the body rows the kernel reads (rigid_state of 2 feet + 2 knees, contact_forces of base + feet; "
                               "identical values, tests/test_env_gpu.py): 488 B per env-step instead of 1,076"),
            "e2e_note": "e2e is OPEN-LOOP: the synthetic simulator states do not depend on the actions, so they are streamed from "
                        "pinned host memory up to 60 steps ahead on a side stream and the actions are not copied back; a real "
                        "simulator on the host would add one D2H of [N,12] actions per step and could not run ahead"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region by a thread of rank 0.  Reads NVML
    in-process (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints): spawning nvidia-smi
    every 200 ms cost 0.4 ms per 9 ms iteration on an 8-GPU box (it enumerates every GPU under the driver lock);
    nvidia-smi remains the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, enabled=True):
        self.index, self.samples, self.stop, self.enabled = index, [], threading.Event(), enabled
        self.t = threading.Thread(target=self.run, daemon=True)
        self.source = None

    def _nvml_open(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = (nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown,
                nv.nvmlClocksEventReasonSwThermalSlowdown, nv.nvmlClocksEventReasonSwPowerCap)

        def sample():
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.samples.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] +
                                ["Active" if r & b else "Not Active" for b in bits])
        return sample

    def run(self):
        if self._sample is not None:
            while not self.stop.is_set():
                self._sample()
                self.stop.wait(0.02)
            return
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.5)

    def __enter__(self):
        self._sample = None
        if os.environ.get("HG_NO_CLOCK_SAMPLER") or not self.enabled:      # env switch: experiments only, never for a reported number
            return self
        try:
            self._sample = self._nvml_open()      # NVML is initialised BEFORE the timed region starts
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.t.is_alive():
            self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


# ------------------------------------------------------------------------------------------ reference / CPU baseline
# Two arms of the reference's algorithm, both test / baseline infrastructure (never on the product path):
#   kind "reference": the UNMODIFIED reference -- humanoid/algo/ppo (PPO, ActorCritic, RolloutStorage) and the text of
#                     humanoid/envs/custom/humanoid_env.py under the stub isaacgym of oracle/ref_harness.py -- loaded
#                     from oracle/_ref/ (recipe oracle/build_ref.py; byte-identical copies, sha256 in its MANIFEST.json),
#   kind "port":      the oracle restatement (oracle/{ppo,env}_oracle.py, pinned bit-exact to the reference), used only
#                     when oracle/_ref/ is absent.
# Either runs the loop of OnPolicyRunner.learn (on_policy_runner.py:128-168): T x (act, env tensor math, process_env_step),
# compute_returns, update -- on the host cores (device "cpu") or, same code, on cuda:0 ("reference_torch_cuda").
def _reference_loop(N, T, device):
    """-> iteration() running ONE PPO iteration of the unmodified reference at N envs x T steps on `device`."""
    import torch
    from oracle import env_oracle as eo
    from oracle import ppo_oracle as po
    from oracle import ref_harness as rh
    PPO, ActorCritic, _ = rh.load_learner()
    c = po.xbot_ppo_config()
    spec = eo.xbot_spec(12)
    with torch.device(device):
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            ac = ActorCritic(705, 219, 12, init_noise_std=1.0, **NETS)
        alg = PPO(ac, num_learning_epochs=c.num_learning_epochs, num_mini_batches=c.num_mini_batches, clip_param=c.clip_param,
                  gamma=c.gamma, lam=c.lam, value_loss_coef=c.value_loss_coef, entropy_coef=c.entropy_coef,
                  learning_rate=c.learning_rate, max_grad_norm=c.max_grad_norm, use_clipped_value_loss=True,
                  schedule=c.schedule, desired_kl=c.desired_kl, device=device)
        alg.init_storage(N, T, [705], [219], [12])
        renv = rh.RefEnv(12, N, spec, device=device)
        g = torch.Generator(device=device).manual_seed(0)
        pool = [eo.synthetic_sim_state(spec, N, g) for _ in range(4)]
        out = renv.step(pool[0], torch.randn(N, spec.So, generator=g))
    state = {"obs": out["obs"], "priv": out["priv"], "i": 0}

    def iteration():
        obs, critic_obs = state["obs"], state["priv"]
        with torch.device(device):
            with torch.inference_mode():
                for t in range(T):
                    actions = alg.act(obs, critic_obs)
                    sim = dict(pool[(state["i"] * T + t) % len(pool)])
                    sim["actions"] = torch.clip(actions, -18.0, 18.0)            # step() :633-635
                    out = renv.step(sim, torch.randn(N, spec.So))
                    obs, critic_obs = out["obs"], out["priv"]
                    alg.process_env_step(out["rew"], out["reset"], {"time_outs": out["time_out"]})
                alg.compute_returns(critic_obs)
            alg.update()
        state["obs"], state["priv"] = obs, critic_obs
        state["i"] += 1
    return iteration


def _port_loop(N, T, device):
    """-> iteration() of the oracle restatement (fallback when oracle/_ref is absent)."""
    import torch
    from oracle import env_oracle as eo
    from oracle import ppo_oracle as po
    cfg = po.xbot_ppo_config()
    with torch.device(device):
        torch.manual_seed(0)
        p = po.init_params(705, 219, 12, (512, 256, 128), (768, 256, 128), (128, 128), 1.0)
        spec = eo.xbot_spec(12)
        est = eo.EnvState(spec, N)
        g = torch.Generator(device=device).manual_seed(0)
        pool = [eo.synthetic_sim_state(spec, N, g) for _ in range(4)]
        out = eo.env_step(spec, est, pool[0], torch.randn(N, spec.So, generator=g))
    state = {"obs": out["obs"], "priv": out["priv"], "i": 0, "lr": cfg.learning_rate, "adam": po.AdamState()}

    def iteration():
        obs, priv = state["obs"], state["priv"]
        with torch.device(device):
            z = torch.zeros
            ro = po.Rollout(z(T, N, 705), z(T, N, 219), z(T, N, 12), z(T, N, 1), z(T, N, 1, dtype=torch.uint8), z(T, N, 1),
                            z(T, N, 1), z(T, N, 12), z(T, N, 12))
            with torch.no_grad():
                for t in range(T):
                    m = po.policy_mean(p, obs)
                    act = po.sample_actions(m, p["std"], torch.randn(N, 12))
                    v = po.value(p, priv)
                    ro.observations[t], ro.privileged_observations[t], ro.actions[t], ro.values[t] = obs, priv, act, v
                    ro.actions_log_prob[t] = po.log_prob(m, p["std"], act).view(-1, 1)
                    ro.mu[t], ro.sigma[t] = m, m * 0 + p["std"]
                    sim = dict(pool[(state["i"] * T + t) % len(pool)])
                    sim["actions"] = act.clip(-18, 18)
                    out = eo.env_step(spec, est, sim, torch.randn(N, spec.So))
                    obs, priv = out["obs"], out["priv"]
                    ro.rewards[t] = po.bootstrap_rewards(out["rew"], v, out["time_out"], cfg.gamma).view(-1, 1)
                    ro.dones[t] = out["reset"].view(-1, 1)
                ro.returns, ro.advantages = po.gae_returns(ro.rewards, ro.values, ro.dones, po.value(p, priv), cfg.gamma, cfg.lam)
            _, _, _, state["lr"] = po.ppo_update(p, state["adam"], ro, torch.randperm(N * T), cfg, state["lr"])
        state["obs"], state["priv"] = obs, priv
        state["i"] += 1
    return iteration


def reference_kind():
    from oracle import build_ref
    return "reference" if build_ref.available() or os.path.isdir("/root/reference/humanoid/algo/ppo") else "port"


def time_reference(N, T, device="cpu", iters=None, warmup=1, seconds_budget=20.0, kind=None):
    """Times `iters` iterations (or as many as fit `seconds_budget`, at least 1) of the reference's PPO iteration at
    N envs x T steps on `device` after `warmup` untimed ones.  CPU: perf_counter, all host threads; CUDA: CUDA events.
    Returns (cpu_baseline-shaped dict, seconds per iteration)."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    kind = kind or reference_kind()
    iteration = (_reference_loop if kind == "reference" else _port_loop)(N, T, device)
    on_gpu = device != "cpu"
    for _ in range(max(1, warmup)):
        iteration()
    times = []
    if on_gpu:
        torch.cuda.synchronize()
        n = iters or 2
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        evs[0].record()
        for i in range(n):
            iteration()
            evs[i + 1].record()
        torch.cuda.synchronize()
        times = [evs[i].elapsed_time(evs[i + 1]) / 1e3 for i in range(n)]
    else:
        t_end = time.perf_counter() + seconds_budget
        while (iters is None and (not times or time.perf_counter() < t_end) and len(times) < 50) or (iters is not None and len(times) < iters):
            t0 = time.perf_counter()
            iteration()
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    what = ("the UNMODIFIED reference (humanoid/algo/ppo + humanoid_env.py tensor math under a stub isaacgym, from oracle/_ref)"
            if kind == "reference" else "the oracle restatement (pinned bit-exact to the reference on CPU; oracle/_ref absent)")
    where = "torch-CUDA fp32 eager on cuda:0, CUDA-event timed" if on_gpu else "torch-CPU fp32, all host threads"
    return {"value": N * T / med, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "device": device,
            "envs": N, "horizon": T, "ms_per_iteration": med * 1e3,
            "sample": f"{N} envs x {T} steps per iteration (env math + rollout + GAE + update), median of {len(times)} "
                      f"iterations after {max(1, warmup)} warm-up, {med * 1e3:.0f} ms/iteration; {what}; {where}"}, med


def cpu_baseline(a, seconds_budget=20.0):
    """Reported CPU baseline inside the product line: a bounded sample (`--cpu-sample-envs` envs x horizon) on the host."""
    return time_reference(a.cpu_sample_envs, a.horizon, "cpu", seconds_budget=seconds_budget)


def _reference_arm_envs(a, rate_guess, budget_s=150.0):
    """Envs per iteration for the reference arm so that (steps + warmup) iterations end within `budget_s`: the FULL
    workload when it fits, otherwise the largest power-of-two sample that does (never below 64)."""
    iters = max(1, a.steps) + max(1, a.warmup)
    fit = int(rate_guess * budget_s / (iters * a.horizon))
    if fit >= a.envs:
        return a.envs
    n = 64
    while n * 2 <= fit:
        n *= 2
    return min(n, a.envs)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.ref_envs:
        n = a.ref_envs
    else:      # calibrate the host's rate on one small iteration, then size the sample (CPU throughput is ~flat in N from 256 envs on)
        _, med = time_reference(256, a.horizon, a.ref_device, iters=1, warmup=1)
        n = a.envs if a.ref_device == "cuda" else _reference_arm_envs(a, 256 * a.horizon / med)
    base, med = time_reference(n, a.horizon, a.ref_device, iters=max(1, a.steps), warmup=max(1, a.warmup))
    cfg = workload_config(a, world)
    cfg["envs_per_gpu"] = n
    cfg["sampled"] = n != a.envs
    cfg["full_envs_per_gpu"] = a.envs
    cfg["workload"] = cfg["workload"].replace(f"{a.envs} envs/GPU", f"{n} envs" + (f" (bounded sample of the {a.envs}-env workload)" if n != a.envs else ""))
    cfg["parallelism"] = f"host cores x{base['cores']} (rank 0 only)" if a.ref_device == "cpu" else "cuda:0 eager"
    cfg["rng"] = "torch default generator (torch.randn / Normal.sample / randperm), as the reference draws"
    cfg["precision"] = "fp32 throughout (the reference enables neither TF32 nor AMP)"
    cfg["l2_policy"] = "n/a (host arm)" if a.ref_device == "cpu" else "working set exceeds L2"
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": med * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "cpu_baseline": base,
            "ref_device": a.ref_device,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import torch
    import torch.distributed as dist
    from humanoid_gym_with_comments_b200 import _lib
    from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
    from humanoid_gym_with_comments_b200.algo.ppo import engine as eng_mod
    from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the JSON line only (NCCL_DEBUG=VERSION prints there)
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = _lib.load()
    N, T = a.envs, a.horizon

    torch.manual_seed(0)                         # identical initial weights on every rank
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, 12, init_noise_std=1.0, **NETS)
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
              entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive",
              desired_kl=0.01, device=dev)
    alg.init_storage(N, T, [705], [219], [12])
    alg.torch_rng = a.torch_rng
    torch.manual_seed(1000 + rank)               # per-rank env / sampling streams
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), dev, pool=4, seed=rank, torch_rng_noise=a.torch_rng,
                           compact_bodies=not a.full_body_tables)
    if not a.unbound_env:
        env.bind_storage(alg.storage, gamma=alg.gamma, ring=not a.stacked)
    obs, critic_obs = env.get_observations(), env.get_privileged_observations()
    state = {"obs": obs, "critic_obs": critic_obs}
    h2d_bytes_per_env_step = sum(v.numel() * v.element_size() for v in env.pool[0].values())

    def iteration(e2e: bool):
        obs, critic_obs = state["obs"], state["critic_obs"]
        with torch.inference_mode():
            with _lib.nvtx_range("hg.rollout"):
                for i in range(T):
                    actions = alg.act(obs, critic_obs)
                    obs, privileged_obs, rewards, dones, infos = env.step(actions)
                    critic_obs = privileged_obs if privileged_obs is not None else obs
                    alg.process_env_step(rewards, dones, infos)
            alg.compute_returns(critic_obs)
        losses = alg.update()                    # reads the loss log back (D2H) once per iteration
        state["obs"], state["critic_obs"] = obs, critic_obs
        if e2e:                                  # + the iteration's mean (bootstrapped) reward, one reduction over [T, N]
            return losses, float(alg.storage.rewards.mean().item())
        return losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, e2e):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.hg_launch_count()
        ev0.record()
        for _ in range(k):
            iteration(e2e)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, lib.hg_launch_count() - l0

    for _ in range(max(3, a.warmup)):
        iteration(False)
    with ClockSampler(local, enabled=(rank == 0)) as clocks:      # one sampler per job: N concurrent nvidia-smi pollers perturb the launches
        ms, launches = timed(a.steps, False)
    value = N * T * world * a.steps / (ms / 1e3)
    allreduce_kind = getattr(alg, "allreduce_kind", "none")
    e2e = None
    if not a.no_e2e:
        env.stream_states_from_host(True)    # every step's state: pinned host -> device (side stream, 12-step chunks, ring of 6)
        iteration(True)
        ms_e, _ = timed(a.steps, True)
        e2e = {"value": N * T * world * a.steps / (ms_e / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": env.h2d_bytes_per_step * T, "d2h_bytes_per_step": 8 * 8 * 8 + 4,
               "ms_per_step": ms_e / a.steps}

    env.stream_states_from_host(False)

    # ---- roofline of the dominant kernel: every grouped tcgen05 GEMM launch of one more iteration, CUDA-event timed
    records = []
    orig = eng_mod.MLPEngine._launch

    def probe(self, problems, epi, block_n):
        fl = sum(2.0 * p.M * p.N * p.K for p in problems)
        # algorithmic HBM bytes of the launch as the layers are laid out (every distinct bf16 operand read once, the
        # output written once; DELU also reads the activation tile, the split-K RED output is one fp32 write)
        seen, by = set(), 0.0
        for p in problems:
            for key, nbytes in (((p.A, p.M, p.K), 2.0 * p.M * p.K), ((p.B, p.N, p.K), 2.0 * p.N * p.K)):
                if key not in seen:
                    seen.add(key)
                    by += nbytes
            by += {0: 4.0, 1: 2.0, 2: 4.0, 3: 4.0, 4: 2.0}[epi] * p.M * p.N
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(self, problems, epi, block_n)
        e1.record()
        records.append((epi, block_n, fl, e0, e1, [(p.M, p.N, p.K) for p in problems], by))

    eng_mod.MLPEngine._launch = probe
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    iteration(False)
    ev1.record()
    torch.cuda.synchronize()
    eng_mod.MLPEngine._launch = orig
    iter_ms = ev0.elapsed_time(ev1)
    by_kind = {}
    for epi, bn, fl, e0, e1, shapes, by in records:
        big = not (shapes[0][0] == N or shapes[0][2] == N)      # rollout launches have N rows, update launches N*T/4
        key = f"{['store','fwd_bias_elu','dx_delu','dw_atomic','bias'][epi]}_bn{bn}_{'update' if big else 'rollout'}"
        d = by_kind.setdefault(key, {"launches": 0, "ms": 0.0, "flop": 0.0, "bytes": 0.0})
        d["launches"] += 1
        d["ms"] += e0.elapsed_time(e1)
        d["flop"] += fl
        d["bytes"] += by
    upd = {k: v for k, v in by_kind.items() if k.endswith("update")}
    g_ms = sum(v["ms"] for v in upd.values())
    g_fl = sum(v["flop"] for v in upd.values())
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_BF16_TFLOPS))
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
    g_by = sum(v["bytes"] for v in upd.values())
    # the layer-by-layer design sits at the ridge (~190 FLOP/B against 211): per kind, the floor is the slower of the two roofs
    floor_ms = sum(max(v["flop"] / (peak * 1e12), v["bytes"] / (hbm_peak * 1e9)) * 1e3 for v in upd.values())
    achieved = g_fl / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:      # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture (never measured under ncu here)
        name = next(n for n in ("r02_ncu_update_gemms.json", "r01_ncu_update_gemms.json") if os.path.exists(os.path.join(ROOT, "profiles", n)))
        tj = json.load(open(os.path.join(ROOT, "profiles", name)))
        traffic, traffic_src = float(tj["traffic_bytes_per_launch"]), f"profiles/{name} (dram read + write, mean over one minibatch's 8 launches)"
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "hg::gemm_kernel (grouped tcgen05 GEMM; update-phase launches)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                "traffic": traffic, "traffic_source": traffic_src,
                "flop_per_launch": g_fl / max(1, sum(v["launches"] for v in upd.values())),
                "launches": sum(v["launches"] for v in upd.values()), "kernel_ms_per_iteration": g_ms,
                "share_of_iteration": g_ms / iter_ms if iter_ms else None,
                "hbm_view": {"algorithmic_bytes_per_launch": g_by / max(1, sum(v["launches"] for v in upd.values())),
                             "achieved_gbs": g_by / (g_ms / 1e3) / 1e9 if g_ms > 0 else None, "peak_gbs": hbm_peak,
                             "frac": g_by / (g_ms / 1e3) / 1e9 / hbm_peak if g_ms > 0 else None,
                             "frac_of_binding_roof": floor_ms / g_ms if g_ms > 0 else None,
                             "note": "same launches against the HBM roof: activations round-trip through HBM between the "
                                     "layer launches, so the short-K launches (dX, critic layer 1, layer 3) are bound by "
                                     "bytes, not FLOPs; frac_of_binding_roof = sum over kinds of max(tensor floor, HBM floor) / measured"},
                "by_kind": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                                "tflops": round(v["flop"] / (v["ms"] / 1e3) / 1e12, 2) if v["ms"] > 0 else None,
                                "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["ms"] > 0 else None}
                            for k, v in sorted(by_kind.items())}}

    # ---- the HBM-bound kernels of the path against the measured copy bandwidth (SURVEY 8d figures), CUDA-event timed
    def time_calls(fn, reps):
        fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps        # us per call

    aux = {}
    try:
        st = alg.storage
        lastv = torch.zeros(N, 1, device=dev)
        us = time_calls(lambda: st.compute_returns(lastv, alg.gamma, alg.lam), 20)
        gae_bytes = 25.0 * N * T                     # scan 17 B + normalisation pass 8 B per (t, env)
        aux["gae_returns"] = {"us": us, "bytes": gae_bytes, "gbs": gae_bytes / us / 1e3, "frac": gae_bytes / us / 1e3 / hbm_peak,
                              "launches_per_call": 2}
        acts = torch.zeros(N, 12, device=dev)
        slot_i = [0]

        def env_once():                              # walks through the storage slots like the rollout does
            st.step = slot_i[0] % (T - 1)
            slot_i[0] += 1
            env.step(acts)
            st.env_stored = None
        us = time_calls(env_once, min(T - 1, 40))
        if st.ring_enabled:      # history ring: newest frame only -- fp32 live + fp32 archive + bf16 live, actor and critic, + ~0.9 KB state
            env_bytes = (2 * 4.0 * (47 + 73) + 2 * 2.0 * (48 + 80) + 900.0) * N     # fp32 live + archive, bf16 live + archive, state
        else:                    # SURVEY 8d K-a: history shift + state, D = 12
            env_bytes = (4.0 * (47 * 29 + 73 * 5) + 900.0) * N
        st.step = 0
        aux["env_step"] = {"us": us, "bytes": env_bytes, "gbs": env_bytes / us / 1e3, "frac": env_bytes / us / 1e3 / hbm_peak,
                           "launches_per_call": 1, "format": "history ring" if st.ring_enabled else "stacked"}
    except Exception as e:       # never lose the headline line to an auxiliary measurement
        aux["error"] = repr(e)

    base, ref_cuda = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        base, _ = cpu_baseline(a)
    if rank == 0 and world == 1 and not a.no_ref_cuda:
        # the denominator of north_star's ">= 50x the reference's torch-CUDA PPO iteration": the unmodified reference on this
        # very GPU at the FULL workload (eager torch, fp32), 1 warm-up + 2 CUDA-event-timed iterations, after our own timing
        try:
            rc, _ = time_reference(N, T, dev, iters=2, warmup=1)
            rc["speedup_of_value"] = value / rc["value"]
            if e2e:
                rc["speedup_of_e2e"] = e2e["value"] / rc["value"]
            ref_cuda = rc
        except Exception as e:
            ref_cuda = {"error": repr(e)}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": dict(workload_config(a, world), collective=(
                    "none (one GPU)" if world == 1 else
                    {"symm_multimem": "gradient arena in symmetric memory, two-shot multimem (NVSwitch in-switch reduce) all-reduce of 4.13 MB per minibatch",
                     "symm_two_shot": "gradient arena in symmetric memory, peer-to-peer two-shot all-reduce of 4.13 MB per minibatch",
                     "nccl": "ncclAllReduce of 4.13 MB per minibatch"}.get(allreduce_kind, allreduce_kind))),
                "clocks": clocks.summary(), "e2e": e2e,
                "gpu_launches": int(launches), "gpu_launches_per_step": launches / a.steps,
                "roofline": roofline, "hbm_kernels": aux, "cpu_baseline": base, "reference_torch_cuda": ref_cuda}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly ONE line, the JSON: anything a library prints there meanwhile (NCCL's version banner comes from
    # C code on rank 0) is sent to stderr by pointing fd 1 at fd 2 until the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if a.impl == "reference":
                run_reference(a)
            else:
                run_ours(a)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
        other = [l for l in buf.getvalue().splitlines() if l and not l.startswith("{")]
        if other:
            print("\n".join(other), file=sys.stderr)
        if lines:
            os.write(real_stdout, (lines[-1] + "\n").encode())
        os.close(real_stdout)


if __name__ == "__main__":
    main()
