"""Env-side hot path: the tensor math of `XBotLFreeEnv.post_physics_step` as ONE fused CUDA launch.

Reference being replaced (humanoid/envs/custom/humanoid_env.py): base-frame quantities :784-788,
`check_termination` :811-816, `compute_reward` + 22 `_reward_*` :889-907/:1170-1437, the buffer and history
zeroing of `reset_idx` :1132-1139/:1160-1163, `compute_observations` (+ `compute_ref_state`, `_get_gait_phase`,
noise, 15/3-frame history) :683-744/:818-887, the hand-over :802-806 and the +-18 clips :654-657.

`XBotEnvMath` keeps the reference's attribute names for the persistent tensors so a subclass of the real
`XBotLFreeEnv` can delegate to it.  `SyntheticXBotEnv` is the `VecEnv`-shaped stand-in used by
bench.py and the tests: Isaac Gym's closed PhysX step is replaced by a pool of synthetic Isaac-Gym-shaped state
tensors.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from .. import _lib
from .._lib import (CMD_CALLBACK, CMD_RESET, ENV_PHASE_ALL, ENV_PHASE_OBS, ENV_PHASE_REWARD, HG_NUM_REWARDS, REWARD_NAMES,
                    HgActionArgs, HgCommandArgs, HgEnvBuffers, HgEnvConfig, check, current_stream, ptr)

# humanoid_config.py:395-430 (XBotLCfg.rewards.scales)
XBOT_REWARD_SCALES = dict(
    joint_pos=1.6, feet_clearance=1.0, feet_contact_number=1.2, feet_air_time=1.0, foot_slip=-0.05,
    feet_distance=0.2, knee_distance=0.2, feet_contact_forces=-0.01, tracking_lin_vel=1.2,
    tracking_ang_vel=1.1, vel_mismatch_exp=0.5, low_speed=0.2, track_vel_hard=0.5,
    default_joint_pos=0.5, orientation=1.0, base_height=0.2, base_acc=0.2, action_smoothness=-0.002,
    torques=-1e-5, dof_vel=-5e-4, dof_acc=-1e-7, collision=-1.0)


@dataclass
class XBotMathConfig:
    """The scalars of XBotLCfg that the env math reads (humanoid_config.py: env :37-68, noise :162-182,
    rewards :372-430, normalization :432-459) + the body/DOF layout."""
    num_envs: int = 4096
    num_dof: int = 12
    leg_offset: int = 0                  # 0: upstream 12-DOF XBot-L, 6: shipped 18-DOF fork (arms first)
    num_bodies: int = 13
    feet_indices: tuple = (6, 12)
    knee_indices: tuple = (4, 10)
    termination_contact_indices: tuple = (0,)
    penalised_contact_indices: tuple = (0,)
    frame_stack: int = 15
    c_frame_stack: int = 3
    dt: float = 0.01
    cycle_time: float = 0.64
    episode_length_s: float = 24.0
    target_joint_pos_scale: float = 0.17
    base_height_target: float = 0.94
    min_dist: float = 0.2
    max_dist: float = 0.5
    target_feet_height: float = 0.1
    tracking_sigma: float = 5.0
    max_contact_force: float = 700.0
    only_positive_rewards: bool = True
    lin_vel_scale: float = 2.0
    ang_vel_scale: float = 1.0
    dof_pos_scale: float = 1.0
    dof_vel_scale: float = 0.05
    quat_scale: float = 1.0
    clip_observations: float = 18.0
    add_noise: bool = True
    noise_level: float = 0.6
    noise_dof_pos: float = 0.05
    noise_dof_vel: float = 0.5
    noise_ang_vel: float = 0.1
    noise_quat: float = 0.03
    reward_scales: Dict[str, float] = field(default_factory=lambda: dict(XBOT_REWARD_SCALES))
    phase_mul_reciprocal: bool = False   # True reproduces torch-CUDA's x/c == x*(1/c)
    # commands / domain_rand scalars read by _post_physics_step_callback (humanoid_config.py:331-337,349-370)
    resampling_time: float = 8.0
    heading_command: bool = True
    command_ranges: Dict[str, tuple] = field(default_factory=lambda: dict(
        lin_vel_x=(-0.3, 0.6), lin_vel_y=(-0.3, 0.3), ang_vel_yaw=(-0.3, 0.3), heading=(-3.14, 3.14)))
    push_robots: bool = True
    push_interval_s: float = 4.0
    max_push_vel_xy: float = 0.2
    max_push_ang_vel: float = 0.4

    @staticmethod
    def for_dof(num_dof: int, num_envs: int = 4096) -> "XBotMathConfig":
        if num_dof == 12:
            return XBotMathConfig(num_envs=num_envs)
        if num_dof == 18:
            return XBotMathConfig(num_envs=num_envs, num_dof=18, leg_offset=6, num_bodies=19,
                                  feet_indices=(12, 18), knee_indices=(10, 16))
        raise ValueError("num_dof must be 12 or 18")

    @property
    def num_single_obs(self) -> int:
        return 11 + 3 * self.num_dof

    @property
    def single_num_privileged_obs(self) -> int:
        return 25 + 4 * self.num_dof

    @property
    def num_observations(self) -> int:
        return self.frame_stack * self.num_single_obs

    @property
    def num_privileged_obs(self) -> int:
        return self.c_frame_stack * self.single_num_privileged_obs

    def noise_scale_vec(self) -> torch.Tensor:
        """`_get_noise_scale_vec` (:748-768).  The 18-DOF fork's hard-coded slices (misaligned with its
        65-wide frame) are reproduced as data."""
        v = torch.zeros(self.num_single_obs)
        D = self.num_dof
        if D == 18:
            v[5:23] = self.noise_dof_pos * self.dof_pos_scale
            v[23:41] = self.noise_dof_vel * self.dof_vel_scale
            v[53:56] = self.noise_ang_vel * self.ang_vel_scale
            v[56:59] = self.noise_quat * self.quat_scale
        else:
            v[5:5 + D] = self.noise_dof_pos * self.dof_pos_scale
            v[5 + D:5 + 2 * D] = self.noise_dof_vel * self.dof_vel_scale
            v[5 + 3 * D:8 + 3 * D] = self.noise_ang_vel * self.ang_vel_scale
            v[8 + 3 * D:11 + 3 * D] = self.noise_quat * self.quat_scale
        return v


class XBotEnvMath:
    """Persistent env-math state + the fused step.  Attribute names follow `XBotLFreeEnv._init_buffers`."""

    def __init__(self, cfg: XBotMathConfig, device, default_dof_pos: Optional[torch.Tensor] = None,
                 noise_scale_vec: Optional[torch.Tensor] = None):
        if torch.device(device).type != "cuda":
            raise _lib.HgError("XBotEnvMath needs a CUDA device (no CPU fallback)")
        for k in cfg.reward_scales:
            if k not in REWARD_NAMES and cfg.reward_scales[k] != 0:
                raise NotImplementedError(f"reward term '{k}' has no fused implementation")
        self.cfg, self.device = cfg, torch.device(device)
        N, D = cfg.num_envs, cfg.num_dof
        z = lambda *s, **k: torch.zeros(*s, device=self.device, **k)
        self.num_envs, self.num_dof = N, D
        self.max_episode_length = math.ceil(cfg.episode_length_s / cfg.dt)
        self.actions, self.last_actions, self.last_last_actions = z(N, D), z(N, D), z(N, D)
        self.last_dof_vel, self.last_root_vel = z(N, D), z(N, 6)
        self.feet_air_time = z(N, 2)
        self.last_contacts = z(N, 2, dtype=torch.bool)
        self.feet_height = z(N, 2)
        self.last_feet_z = torch.full((N, 2), 0.05, device=self.device)
        self.ref_dof_pos = z(N, D)
        self.episode_length_buf = z(N, dtype=torch.long)
        self.episode_sums_tensor = z(HG_NUM_REWARDS, N)
        self.default_dof_pos = z(1, D) if default_dof_pos is None else default_dof_pos.reshape(1, D).to(self.device).float().contiguous()
        nv = cfg.noise_scale_vec() if noise_scale_vec is None else noise_scale_vec
        self.noise_scale_vec = nv.to(self.device).float().contiguous()
        self.rew_buf = z(N)
        self.reset_buf = z(N, dtype=torch.bool)
        self.time_out_buf = z(N, dtype=torch.bool)
        self.reward_terms = None
        # double-buffered stacked observations (history in -> out must not alias)
        self._obs = [z(N, cfg.num_observations), z(N, cfg.num_observations)]
        self._priv = [z(N, cfg.num_privileged_obs), z(N, cfg.num_privileged_obs)]
        self._cur = 0
        self.obs_buf, self.privileged_obs_buf = self._obs[0], self._priv[0]
        self.step_counter = 0
        self.seed = 0

        c = HgEnvConfig()
        c.num_envs, c.D, c.leg_offset, c.num_bodies = N, D, cfg.leg_offset, cfg.num_bodies
        c.feet[0], c.feet[1] = cfg.feet_indices
        c.knees[0], c.knees[1] = cfg.knee_indices
        c.num_term_bodies = len(cfg.termination_contact_indices)
        for i, b in enumerate(cfg.termination_contact_indices):
            c.term_bodies[i] = b
        c.num_pen_bodies = len(cfg.penalised_contact_indices)
        for i, b in enumerate(cfg.penalised_contact_indices):
            c.pen_bodies[i] = b
        c.frame_stack, c.c_frame_stack = cfg.frame_stack, cfg.c_frame_stack
        c.dt, c.cycle_time, c.max_episode_length = cfg.dt, cfg.cycle_time, float(self.max_episode_length)
        c.target_joint_pos_scale, c.base_height_target = cfg.target_joint_pos_scale, cfg.base_height_target
        c.min_dist, c.max_dist, c.target_feet_height = cfg.min_dist, cfg.max_dist, cfg.target_feet_height
        c.tracking_sigma, c.max_contact_force = cfg.tracking_sigma, cfg.max_contact_force
        c.only_positive_rewards = 1 if cfg.only_positive_rewards else 0
        c.lin_vel_scale, c.ang_vel_scale, c.dof_pos_scale = cfg.lin_vel_scale, cfg.ang_vel_scale, cfg.dof_pos_scale
        c.dof_vel_scale, c.quat_scale, c.clip_obs = cfg.dof_vel_scale, cfg.quat_scale, cfg.clip_observations
        c.noise_level = cfg.noise_level if cfg.add_noise else 0.0
        for k, name in enumerate(REWARD_NAMES):
            c.reward_scale_dt[k] = cfg.reward_scales.get(name, 0.0) * cfg.dt     # python double product, as :212
        c.phase_mul_reciprocal = 1 if cfg.phase_mul_reciprocal else 0
        c.phases = ENV_PHASE_ALL
        self._c = c
        self._buf_cache = {}
        self._lib = _lib.load()

    # ------------------------------------------------------------------ callback, pushes, extras
    def _callback_state(self):
        if not hasattr(self, "commands"):
            with torch.inference_mode(False):          # persistent buffers: never inference tensors, whoever triggers this first
                self._make_callback_state()

    def _make_callback_state(self):
        N, z = self.num_envs, lambda *s, **k: torch.zeros(*s, device=self.device, **k)
        self.commands = z(N, 4)                      # x vel, y vel, yaw vel, heading (:271)
        self.rand_push_force, self.rand_push_torque = z(N, 3), z(N, 3)
        self.root_vel_prepush = z(N, 6)
        self.common_step_counter = 0
        self.episode_means_tensor = z(HG_NUM_REWARDS)
        self.episode_reset_count = z(1, dtype=torch.int32)
        cfg = self.cfg
        self.resampling_steps = int(cfg.resampling_time / cfg.dt)                  # :1005
        self.push_interval = math.ceil(cfg.push_interval_s / cfg.dt)               # :191
        a = HgCommandArgs()
        a.num_envs, a.resampling_steps, a.heading_command = N, self.resampling_steps, 1 if cfg.heading_command else 0
        for i, k in enumerate(("lin_vel_x", "lin_vel_y", "ang_vel_yaw", "heading")):
            lo, hi = cfg.command_ranges[k]
            a.cmd_scale[i], a.cmd_lo[i] = hi - lo, lo                              # python double difference, cast once
        a.push_vel_scale, a.push_vel_lo = cfg.max_push_vel_xy - (-cfg.max_push_vel_xy), -cfg.max_push_vel_xy
        a.push_ang_scale, a.push_ang_lo = cfg.max_push_ang_vel - (-cfg.max_push_ang_vel), -cfg.max_push_ang_vel
        a.commands, a.rand_push_force, a.rand_push_torque = ptr(self.commands), ptr(self.rand_push_force), ptr(self.rand_push_torque)
        a.root_vel_prepush = ptr(self.root_vel_prepush)
        a.episode_length_buf, a.reset_buf = ptr(self.episode_length_buf), ptr(self.reset_buf)
        a.seed = self.seed
        self._cmd_args = a

    def post_physics_callback(self, root_states: torch.Tensor, u_cmd: Optional[torch.Tensor] = None,
                              u_push: Optional[torch.Tensor] = None, ep_increment: int = 1) -> bool:
        """`_post_physics_step_callback` (humanoid_env.py:1001-1016), one launch, BEFORE the reward kernel: periodic
        command resampling, heading -> yaw-rate command, and every `push_interval` steps the random push written into
        `root_states` (in place; the pre-push velocities go to `self.root_vel_prepush`, which `step` must then be given).
        `ep_increment=1` because the env kernel performs `episode_length_buf += 1` (:780) itself.
        u_cmd [N,3] / u_push [N,5]: torch.rand draws, or None for the in-kernel Philox.  Returns whether it pushed."""
        self._callback_state()
        _lib.require_cuda(root_states, u_cmd, u_push)
        self.common_step_counter += 1                                                  # :781
        a = self._cmd_args
        a.mode, a.ep_increment = CMD_CALLBACK, ep_increment
        pushed = bool(self.cfg.push_robots and self.common_step_counter % self.push_interval == 0)
        a.push = 1 if pushed else 0
        a.root_states, a.u_cmd, a.u_push = ptr(root_states), ptr(u_cmd), ptr(u_push)
        a.counter = self.common_step_counter
        check(self._lib.hg_command_step(C.byref(a), current_stream()))
        return pushed

    def reset_bookkeeping(self, u_cmd: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """The parts of `reset_idx` that are neither simulator calls nor buffer zeroing, two launches, BETWEEN the
        HG_ENV_PHASE_REWARD and HG_ENV_PHASE_OBS launches: `_resample_commands(env_ids)` (:1127) and
        extras["episode"]["rew_*"] (:1141-1143).  Returns the extras dict (device scalars, views of one [22] tensor;
        left unchanged on steps without a reset, as the reference's early return does)."""
        self._callback_state()
        _lib.require_cuda(u_cmd)
        a = self._cmd_args
        a.mode, a.push, a.u_cmd, a.u_push = CMD_RESET, 0, ptr(u_cmd), 0
        a.counter = self.common_step_counter
        check(self._lib.hg_command_step(C.byref(a), current_stream()))
        check(self._lib.hg_episode_means(ptr(self.episode_sums_tensor), ptr(self.reset_buf), HG_NUM_REWARDS, self.num_envs,
                                         float(self.cfg.episode_length_s), ptr(self.episode_means_tensor),
                                         ptr(self.episode_reset_count), current_stream()))
        return {"rew_" + name: self.episode_means_tensor[k] for k, name in enumerate(REWARD_NAMES)
                if self.cfg.reward_scales.get(name, 0.0) != 0}

    def step_with_callbacks(self, sim: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None,
                            u_cmd: Optional[torch.Tensor] = None, u_push: Optional[torch.Tensor] = None,
                            u_cmd_reset: Optional[torch.Tensor] = None, **kw):
        """The whole tensor side of `post_physics_step` (:770-806) in the reference's order, five launches:
        callback -> termination + rewards -> reset bookkeeping (commands, extras) -> observations + history + hand-over.
        `sim["root_states"]` is modified in place by a push; commands / push buffers are this object's persistent ones.
        Returns (obs, priv, rew, reset, time_out, extras_episode)."""
        self._callback_state()
        pushed = self.post_physics_callback(sim["root_states"], u_cmd, u_push)
        sim = dict(sim, commands=self.commands, rand_push_force=self.rand_push_force, rand_push_torque=self.rand_push_torque)
        pre = self.root_vel_prepush if pushed else None
        self.step(sim, noise, phases=ENV_PHASE_REWARD, root_vel_prepush=pre, **kw)
        episode = self.reset_bookkeeping(u_cmd_reset)
        sim.pop("actions", None)                       # already copied into self.actions by the first launch
        kw.pop("store", None)                          # the transition record belongs to the reward launch
        out = self.step(sim, noise, phases=ENV_PHASE_OBS, root_vel_prepush=pre, **kw)
        return (*out, episode)

    @property
    def episode_sums(self) -> Dict[str, torch.Tensor]:
        return {name: self.episode_sums_tensor[k] for k, name in enumerate(REWARD_NAMES)
                if self.cfg.reward_scales.get(name, 0.0) != 0}

    def reset_history(self):
        for b in self._obs + self._priv:
            b.zero_()

    def pre_physics(self, raw_actions: Optional[torch.Tensor], dof_state: Optional[torch.Tensor] = None,
                    p_gains: Optional[torch.Tensor] = None, d_gains: Optional[torch.Tensor] = None,
                    torque_limits: Optional[torch.Tensor] = None, torques: Optional[torch.Tensor] = None,
                    delay: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                    dyn_rand: float = 0.02, clip_actions: float = 18.0, action_scale: float = 0.25):
        """The tensor math in front of the physics step, one launch: action delay mix +
        multiplicative noise + clip into `self.actions` (`step`, humanoid_env.py:624-636) and, if `torques` is given,
        the PD law of `_compute_torques` (:910-925).  `raw_actions=None` re-evaluates only the torques (what a real env
        does on each of its physics sub-steps).  `delay` [N] / `noise` [N,D]: torch draws, or None for Philox."""
        a = HgActionArgs()
        a.num_envs, a.D = self.num_envs, self.num_dof
        a.raw_actions, a.delay, a.noise = ptr(raw_actions), ptr(delay), ptr(noise)
        self.step_counter += 1
        a.seed, a.counter = self.seed, self.step_counter
        a.dyn_rand, a.clip_actions, a.action_scale = dyn_rand, clip_actions, action_scale
        a.actions = ptr(self.actions)
        a.dof_state, a.p_gains, a.d_gains = ptr(dof_state), ptr(p_gains), ptr(d_gains)
        a.default_dof_pos, a.torque_limits, a.torques = ptr(self.default_dof_pos), ptr(torque_limits), ptr(torques)
        _lib.require_cuda(raw_actions, dof_state, p_gains, d_gains, torque_limits, torques, delay, noise)
        check(self._lib.hg_action_torque(C.byref(a), current_stream()))
        return self.actions, torques

    def _buffers_for(self, sim: Dict[str, torch.Tensor]) -> HgEnvBuffers:
        """HgEnvBuffers with every static pointer filled in, cached per set of sim tensors."""
        key = tuple(sim[k].data_ptr() for k in _SIM_KEYS)
        b = self._buf_cache.get(key)
        if b is None:
            b = HgEnvBuffers()
            for k in _SIM_KEYS:
                t = sim[k]
                if not t.is_cuda:
                    raise _lib.HgError(f"sim['{k}'] must live on the CUDA device (no CPU fallback)")
                if t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.HgError(f"sim['{k}'] must be a contiguous float32 tensor")
                setattr(b, k, ptr(t))
            b.default_dof_pos, b.noise_scale_vec = ptr(self.default_dof_pos), ptr(self.noise_scale_vec)
            for k in ("actions", "last_actions", "last_last_actions", "last_dof_vel", "last_root_vel", "feet_air_time",
                      "last_contacts", "feet_height", "last_feet_z", "ref_dof_pos", "episode_length_buf"):
                setattr(b, k, ptr(getattr(self, k)))
            b.episode_sums = ptr(self.episode_sums_tensor)
            b.rew_buf, b.reset_buf, b.time_out_buf = ptr(self.rew_buf), ptr(self.reset_buf), ptr(self.time_out_buf)
            b.seed = self.seed
            if len(self._buf_cache) > 64:
                self._buf_cache.clear()
            self._buf_cache[key] = b
        return b

    def step(self, sim: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None, phases: int = ENV_PHASE_ALL,
             obs_out: Optional[torch.Tensor] = None, priv_out: Optional[torch.Tensor] = None,
             obs16_out: Optional[torch.Tensor] = None, priv16_out: Optional[torch.Tensor] = None,
             raw_actions: Optional[torch.Tensor] = None, clip_actions: float = 18.0, store=None,
             root_vel_prepush: Optional[torch.Tensor] = None, ring=None):
        """One fused post-physics step on the Isaac-Gym-shaped tensors in `sim` (keys: root_states, dof_state,
        contact_forces, rigid_state, commands, torques, rand_push_force, rand_push_torque, env_frictions, body_mass;
        optional `actions` [N,D] = the actions applied this step, copied into `self.actions` first).
        `noise`: N(0,1) draws [N, num_single_obs] (the reference's `randn_like`, :877) or None for in-kernel Philox.
        `obs_out` / `priv_out` (+ optional bf16 shadows `obs16_out` / `priv16_out`): destinations, e.g. the learner's
        next storage slot, instead of the internal double buffer.
        `store` = (values, rewards_out, dones_out, gamma): also write the learner's transition record
        rewards_out = rew + gamma * values * time_out, dones_out = reset (what PPO.process_env_step would store).
        `ring` = (storage, row): history-ring format -- only the newest frames are written, into the
        learner storage's rings as the last slot of `row`'s window; the returned observations are strided VIEWS of the
        live ring (the reference's obs_buf / privileged_obs_buf, bit for bit), nothing is shifted or copied.
        Returns (obs_buf, privileged_obs_buf, rew_buf, reset_buf, time_out_buf)."""
        acts = sim.get("actions")
        if acts is not None:
            self.actions.copy_(acts)
        nxt = 1 - self._cur
        b = self._buffers_for(sim)
        if ring is not None:
            st, row = ring
            b.ring_obs, b.ring_priv, b.ring_row = st._ring_obs.c, st._ring_priv.c, row
            b.ring_first_valid, b.ring_hist_start = ptr(st.ring_first_valid), ptr(st.ring_hist_start)
            obs_next, priv_next = st._ring_obs.view32(row), st._ring_priv.view32(row)
            obs_out = obs_next                      # (not the internal double buffer: `_cur` stays)
            if phases & 2:
                st._ring_version += 1
        else:
            b.ring_obs.live32 = None
            obs_next = obs_out if obs_out is not None else self._obs[nxt]
            priv_next = priv_out if priv_out is not None else self._priv[nxt]
        b.noise = ptr(noise)
        b.actions_in, b.clip_actions = ptr(raw_actions), clip_actions   # fused clip of step() (:633-635)
        self.step_counter += 1
        b.counter = self.step_counter
        if ring is None:
            b.obs_prev, b.obs_prev_ld = ptr(self.obs_buf), self.obs_buf.stride(0)
            b.priv_prev, b.priv_prev_ld = ptr(self.privileged_obs_buf), self.privileged_obs_buf.stride(0)
            b.obs_next, b.obs_next_ld = obs_next.data_ptr(), obs_next.stride(0)
            b.priv_next, b.priv_next_ld = priv_next.data_ptr(), priv_next.stride(0)
        if obs16_out is not None:
            b.obs16_next, b.obs16_next_ld = obs16_out.data_ptr(), obs16_out.stride(0)
        else:
            b.obs16_next = 0
        if priv16_out is not None:
            b.priv16_next, b.priv16_next_ld = priv16_out.data_ptr(), priv16_out.stride(0)
        else:
            b.priv16_next = 0
        b.reward_terms = ptr(self.reward_terms)
        if store is not None:
            b.store_values, b.store_rewards, b.store_dones, b.store_gamma = ptr(store[0]), ptr(store[1]), ptr(store[2]), store[3]
        else:
            b.store_values = b.store_rewards = b.store_dones = 0
        b.root_vel_prepush = ptr(root_vel_prepush)     # set on push steps only (post_physics_callback)
        self._c.phases = phases
        check(self._lib.hg_env_step(C.byref(self._c), C.byref(b), current_stream()))
        if phases & 2:
            self.obs_buf, self.privileged_obs_buf = obs_next, priv_next
            if obs_out is None:
                self._cur = nxt
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.time_out_buf


_SIM_KEYS = ("root_states", "dof_state", "contact_forces", "rigid_state", "commands", "torques", "rand_push_force",
             "rand_push_torque", "env_frictions", "body_mass")


def synthetic_sim_state(cfg: XBotMathConfig, gen: torch.Generator, device, upright: bool = True,
                        base_contact_frac: float = 0.02) -> Dict[str, torch.Tensor]:
    """Isaac-Gym-shaped synthetic tensors, generated on `device`."""
    N, D, B = cfg.num_envs, cfg.num_dof, cfg.num_bodies
    rn = lambda *s: torch.randn(*s, generator=gen, device=device)
    ru = lambda *s: torch.rand(*s, generator=gen, device=device)
    root = torch.zeros(N, 13, device=device)
    root[:, :3] = torch.tensor([0.0, 0.0, 0.95], device=device) + 0.1 * rn(N, 3)
    quat = torch.cat((0.1 * rn(N, 3), torch.ones(N, 1, device=device)), dim=1) if upright else rn(N, 4)
    root[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    root[:, 7:13] = 0.3 * rn(N, 6)
    cf = (20.0 * rn(N, B, 3)).abs()
    hit = ru(N) < base_contact_frac
    tb = list(cfg.termination_contact_indices)
    cf[:, tb, :] = torch.where(hit[:, None, None], cf[:, tb, :], 0.05 * rn(N, len(tb), 3))
    rigid = 0.3 * rn(N, B, 13)
    rigid[:, list(cfg.feet_indices), 2] = 0.15 * ru(N, 2)
    return dict(root_states=root, dof_state=0.2 * rn(N * D, 2), contact_forces=cf.contiguous(),
                rigid_state=rigid.contiguous(), commands=0.3 * rn(N, 4), torques=10.0 * rn(N, D),
                rand_push_force=0.2 * rn(N, 3), rand_push_torque=0.4 * rn(N, 3),
                env_frictions=0.1 + 1.9 * ru(N, 1), body_mass=20.0 + 10.0 * ru(N, 1))


class SyntheticXBotEnv:
    """`VecEnv`-shaped (reference humanoid/algo/vec_env.py:39-63) synthetic environment: the PhysX step is
    replaced by cycling through a pool of pre-generated Isaac-Gym-shaped state tensors; everything after it
    (the hot path) is the fused kernel.  Drives `OnPolicyRunner`-style loops in bench.py and the tests."""

    def __init__(self, cfg: XBotMathConfig, device, pool: int = 4, seed: int = 0, torch_rng_noise: bool = True,
                 callbacks: bool = False, compact_bodies: bool = False):
        """callbacks=True also runs `_post_physics_step_callback` (command resampling, heading command, pushes) and the
        command / extras["episode"] part of `reset_idx` every step: five launches per step instead of
        one, commands and push buffers owned by the env instead of taken from the synthetic pool.  Works with
        `bind_storage` like the fused path (the reward launch writes the transition record, the observation launch the
        next storage slot)."""
        self.callbacks = callbacks
        self.cfg, self.device = cfg, torch.device(device)
        self.math = XBotEnvMath(cfg, device)
        self.math.seed = (seed * 0x9E3779B97F4A7C15 + 0x5EED) & (2 ** 64 - 1)     # Philox key of the env's noise: per env instance / rank
        self.num_envs = cfg.num_envs
        self.num_obs = cfg.num_observations
        self.num_privileged_obs = cfg.num_privileged_obs
        self.num_actions = cfg.num_dof
        self.max_episode_length = self.math.max_episode_length
        gen = torch.Generator(device=self.device).manual_seed(seed)
        self.pool = [synthetic_sim_state(cfg, gen, self.device, upright=(i % 2 == 0)) for i in range(pool)]
        if compact_bodies:
            # What a simulator bridge that knows the hot path's inputs would export: of the 13 rigid bodies only the rows the
            # kernel reads -- rigid_state of the two feet and the two knees, contact_forces of the termination / penalised
            # bodies and the feet (humanoid_env.py:811-816,1180-1257,1342-1425).  Same values as the full tables (tested),
            # 488 instead of 1,076 bytes of simulator state per env-step.
            rig = [*cfg.feet_indices, *cfg.knee_indices]
            cfr = sorted({*cfg.termination_contact_indices, *cfg.penalised_contact_indices, *cfg.feet_indices})
            at = {b: i for i, b in enumerate(cfr)}
            for sim in self.pool:
                sim["rigid_state"] = sim["rigid_state"][:, rig].contiguous()
                sim["contact_forces"] = sim["contact_forces"][:, cfr].contiguous()
            c = self.math._c
            c.num_rigid_bodies = 4
            c.rigid_feet[0], c.rigid_feet[1], c.rigid_knees[0], c.rigid_knees[1] = 0, 1, 2, 3
            c.num_bodies = len(cfr)
            c.feet[0], c.feet[1] = at[cfg.feet_indices[0]], at[cfg.feet_indices[1]]
            for i, b in enumerate(cfg.termination_contact_indices):
                c.term_bodies[i] = at[b]
            for i, b in enumerate(cfg.penalised_contact_indices):
                c.pen_bodies[i] = at[b]
        self.math.episode_length_buf.copy_(torch.randint(0, self.max_episode_length, (self.num_envs,), generator=gen,
                                                         device=self.device))
        self.torch_rng_noise = torch_rng_noise
        self.extras = {}
        # OnPolicyRunner.learn reads env.course_gain / env.course_ratio (on_policy_runner.py:160-162); only LeggedRobot
        # defines them (legged_robot.py:84-85), XBotLFreeEnv does not -- a VecEnv that wants to be driven by the
        # unmodified runner has to carry them
        self.course_gain, self.course_ratio = 1.0, 1.0
        self._i = 0
        self._storage = None
        self._ring = False
        self._fast = {}
        self._host_pool = None
        self.obs_buf, self.privileged_obs_buf = self.math.obs_buf, self.math.privileged_obs_buf
        self.rew_buf, self.reset_buf = self.math.rew_buf, self.math.reset_buf
        self.episode_length_buf = self.math.episode_length_buf

    def stream_states_from_host(self, enable: bool = True, chunk: int = 12, chunks: int = 6):
        """End-to-end mode: the Isaac-Gym-shaped state of EVERY step arrives from pinned host memory (as it would from
        an external simulator).  The states of `chunk` consecutive steps travel as ONE copy (53 MB at 4096 envs: 55 GB/s
        on this pool against 38 GB/s for per-step 4.4 MB copies, and a twelfth of the host-side launch work -- per-step
        copies made the rollout host-bound) on a side stream into a ring of `chunks` device staging chunks, so copies
        run up to `chunk * (chunks - 1)` steps ahead of the consumer and also overlap the learner's update.
        `step()` waits on its chunk's copy event when it enters the chunk."""
        if not enable:
            if self._host_pool is not None:
                torch.cuda.current_stream().wait_stream(self._copy_stream)
            self._host_pool = None
            return
        # every state set is ONE packed run of floats (tensors are 256-byte aligned views into it)
        keys = list(self.pool[0].keys())
        offs, total = {}, 0
        for k in keys:
            offs[k] = total
            total += (self.pool[0][k].numel() + 63) // 64 * 64
        self._packed_floats = total
        chunk = max(len(self.pool), chunk // len(self.pool) * len(self.pool))     # whole cycles of the synthetic pool

        def views(buf):
            return {k: buf[offs[k]:offs[k] + self.pool[0][k].numel()].view(self.pool[0][k].shape) for k in keys}

        hb = torch.empty(chunk, total, dtype=torch.float32).pin_memory()          # `chunk` consecutive steps, pinned
        for j in range(chunk):
            for k, v in views(hb[j]).items():
                v.copy_(self.pool[j % len(self.pool)][k])
        self._host_pool = hb
        self._chunk, self._chunks = chunk, chunks
        self._stage_buf = torch.empty(chunks, chunk, total, dtype=torch.float32, device=self.device)
        self._stage = [[views(self._stage_buf[c, j]) for j in range(chunk)] for c in range(chunks)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._copy_done = [torch.cuda.Event() for _ in range(chunks)]
        self._consumed = [torch.cuda.Event() for _ in range(chunks)]
        self.h2d_bytes_per_step = total * 4
        self._i0 = self._i                       # stream position 0 = the next step
        self._next_chunk = 0
        self._prefetch_upto(chunks)

    def _prefetch_upto(self, end_chunk):
        with torch.cuda.stream(self._copy_stream):
            while self._next_chunk < end_chunk:
                slot = self._next_chunk % self._chunks
                self._copy_stream.wait_event(self._consumed[slot])   # the env kernels that last read this chunk have finished
                self._stage_buf[slot].copy_(self._host_pool, non_blocking=True)
                self._copy_done[slot].record(self._copy_stream)
                self._next_chunk += 1

    def bind_storage(self, storage, gamma: Optional[float] = None, ring: bool = True):
        """Let `step()` write the next observations straight into the learner's storage: `PPO.act` then finds them in
        place and launches no ingest / copy (replaces, together with the env kernel, rollout_storage.py:90-91).
        ring=True (default): history-ring format -- the storage keeps ONE copy of every frame
        (`RolloutStorage.enable_ring`), the env writes only the newest frame per step and returns strided views of the
        ring; ring=False: the stacked format (fp32 slot + bf16 shadow of the whole 705 / 219-wide row every step).
        With `gamma` (the learner's discount) the env kernel also stores the transition's bootstrapped reward and done
        flag (ppo.py:129-133), and `PPO.process_env_step` launches nothing when it is handed this env's reward / reset
        buffers.  Pass None to unbind."""
        old = self._storage
        self._fast = {}
        if old is not None and old.ring_enabled and (storage is not old or not ring):
            # hand the current window back to the env's own stacked buffers before leaving the ring
            self.math._obs[self.math._cur].copy_(self.obs_buf)
            self.math._priv[self.math._cur].copy_(self.privileged_obs_buf)
            self.math.obs_buf, self.math.privileged_obs_buf = self.math._obs[self.math._cur], self.math._priv[self.math._cur]
            self.obs_buf, self.privileged_obs_buf = self.math.obs_buf, self.math.privileged_obs_buf
            old.disable_ring()
        self._storage = storage
        self._store_gamma = gamma
        self._ring = bool(ring) and storage is not None
        if self._ring and not storage.ring_enabled:
            cfg = self.cfg
            storage.enable_ring(cfg.num_single_obs, cfg.frame_stack, cfg.single_num_privileged_obs, cfg.c_frame_stack)
            storage.ring_import(self.obs_buf, self.privileged_obs_buf, row=storage.step)     # the env's history so far
            self.obs_buf = self.math.obs_buf = storage._ring_obs.view32(storage.step)
            self.privileged_obs_buf = self.math.privileged_obs_buf = storage._ring_priv.view32(storage.step)

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset(self):
        return self.obs_buf, self.privileged_obs_buf

    def step(self, actions):
        if getattr(self, "_host_pool", None) is not None:
            pos = self._i - self._i0
            c, j = divmod(pos, self._chunk)
            slot = (c % self._chunks, j)
            if j == 0:
                torch.cuda.current_stream().wait_event(self._copy_done[slot[0]])
            sim = self._stage[slot[0]][j]
            sim_id = slot
        else:
            slot = None
            sim_id = self._i % len(self.pool)
            sim = self.pool[sim_id]
        self._i += 1
        if not (actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()):
            actions = actions.to(self.device).float().contiguous()
        noise = torch.randn(self.num_envs, self.cfg.num_single_obs, device=self.device) if (
            self.torch_rng_noise and self.cfg.add_noise) else None
        st = self._storage
        # Bound to a learner's history ring without callbacks: every pointer of the launch is a function of (state set,
        # storage row) only, so the argument block is prepared once per pair and a step is one C call (the generic path
        # below fills ~60 struct fields per step: the rollout would be bound by this host code, not by the GPU).
        fast_key = None
        if self._ring and not self.callbacks and st is not None and getattr(self, "_store_gamma", None) is not None \
                and st.step + 1 <= st.num_transitions_per_env:
            fast_key = (sim_id, st.step)
            rec = self._fast.get(fast_key)
            if rec is not None and rec[0] == actions.data_ptr() and rec[1] is st:
                out = self._step_prepared(rec, noise)
                fast_key = None
            else:
                rec = None
        else:
            rec = None
        if rec is None:
            out = self._step_generic(sim, actions, noise, st, fast_key)
        if slot is not None and slot[1] == self._chunk - 1:     # last step of a chunk: hand its staging chunk back
            self._consumed[slot[0]].record(torch.cuda.current_stream())
            self._prefetch_upto((self._i - self._i0) // self._chunk + self._chunks)
        self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, to = out
        self.extras["time_outs"] = to
        return self.obs_buf, self.privileged_obs_buf, self.rew_buf, self.reset_buf, self.extras

    def _step_prepared(self, rec, noise):
        _, st, b, bref, out, stored = rec
        m = self.math
        m.step_counter += 1
        b.counter = m.step_counter
        b.noise = noise.data_ptr() if noise is not None else 0
        if m._c.phases != ENV_PHASE_ALL:
            m._c.phases = ENV_PHASE_ALL
        if m._lib.hg_env_step(self._cref, bref, current_stream()) != 0:
            check(-1)
        st._ring_version += 1
        st.env_stored = stored
        m.obs_buf, m.privileged_obs_buf = out[0], out[1]
        return out

    def _step_generic(self, sim, actions, noise, st, fast_key=None):
        kw = dict(raw_actions=actions)                  # clip(+-18) of step() (:633-635) fused in the kernel
        if st is not None and st.step + 1 <= st.num_transitions_per_env:
            if self._ring:
                kw.update(ring=(st, st.step + 1))      # one new frame into the learner's history rings
            else:
                o, p, o16, p16 = st.slot(st.step + 1)  # the env writes the learner's next storage slot (+ bf16 shadow) itself
                kw.update(obs_out=o, priv_out=p, obs16_out=o16, priv16_out=p16)
            if getattr(self, "_store_gamma", None) is not None:
                t = st.step
                kw["store"] = (st.values[t], st.rewards[t], st.dones[t], float(self._store_gamma))
                st.env_stored = (t, self.math.rew_buf.data_ptr(), self.math.reset_buf.data_ptr(), float(self._store_gamma))
        if self.callbacks:
            if not hasattr(self, "_root_scratch"):
                with torch.inference_mode(False):      # (a buffer first made under inference_mode could not be written outside it)
                    self._root_scratch = torch.empty(sim["root_states"].shape, device=self.device)
            self._root_scratch.copy_(sim["root_states"])           # a push overwrites root velocities in place
            sim = dict(sim, root_states=self._root_scratch)
            ur = (lambda *sh: torch.rand(*sh, device=self.device)) if self.torch_rng_noise else (lambda *sh: None)
            *out, episode = self.math.step_with_callbacks(sim, noise, u_cmd=ur(self.num_envs, 3), u_push=ur(self.num_envs, 5),
                                                          u_cmd_reset=ur(self.num_envs, 3), **kw)
            self.extras["episode"] = episode
        else:
            out = self.math.step(sim, noise, **kw)
            if fast_key is not None:                   # snapshot of the argument block this call just used
                b = HgEnvBuffers.from_buffer_copy(self.math._buffers_for(sim))
                if len(self._fast) > 4096:
                    self._fast.clear()
                self._cref = C.byref(self.math._c)
                self._fast[fast_key] = (actions.data_ptr(), st, b, C.byref(b), tuple(out), st.env_stored)
        return out
