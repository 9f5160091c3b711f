"""CPU oracle (torch fp32) for the env-side tensor math of the hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Restates `XBotLFreeEnv.post_physics_step` of the reference minus every Isaac Gym call
(`humanoid/envs/custom/humanoid_env.py:770-809`): base-frame quantities (:784-788), termination
(:811-816), the 22 active `_reward_*` terms in alphabetical order (:889-907, :1170-1437), the buffer
and history zeroing of `reset_idx` (:1132-1139, :1160-1163), `compute_observations` with
`compute_ref_state` / `_get_gait_phase` (:683-744, :818-887), the state hand-over (:802-806) and the
+-18 clip of `step` (:654-657).  Parametric in D (12 = upstream XBot-L layout of BASELINE.json,
18 = the shipped fork) and `leg_offset` (0 / 6), as SURVEY.md 0.3 requires.

Pinned by `oracle/make_golden.py` against the unmodified reference run under a stub `isaacgym`
(both layouts; the 12-DOF reference text is obtained by the 13 replacements of SURVEY appendix C).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from .isaac_math import euler_xyz_wrapped, quat_apply, quat_rotate_inverse

# humanoid_config.py:395-430 (XBotLCfg.rewards.scales); zero entries are dropped (:207-212)
XBOT_REWARD_SCALES = dict(
    joint_pos=1.6, feet_clearance=1.0, feet_contact_number=1.2, feet_air_time=1.0, foot_slip=-0.05,
    feet_distance=0.2, knee_distance=0.2, feet_contact_forces=-0.01, tracking_lin_vel=1.2,
    tracking_ang_vel=1.1, vel_mismatch_exp=0.5, low_speed=0.2, track_vel_hard=0.5,
    default_joint_pos=0.5, orientation=1.0, base_height=0.2, base_acc=0.2, action_smoothness=-0.002,
    torques=-1e-5, dof_vel=-5e-4, dof_acc=-1e-7, collision=-1.0, termination=-0.0, feet_stumble=-0.0,
    action_rate=-0.0, stand_still=-0.0)


@dataclass
class EnvSpec:
    """Static description of one env layout + the config scalars the math reads."""
    D: int = 12
    leg_offset: int = 0
    num_bodies: int = 13
    feet: tuple = (6, 12)
    knees: tuple = (4, 10)
    termination_bodies: tuple = (0,)
    penalised_bodies: tuple = (0,)
    frame_stack: int = 15
    c_frame_stack: int = 3
    dt: float = 0.01
    cycle_time: float = 0.64
    max_episode_length: float = 2400.0
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
    clip_obs: float = 18.0
    add_noise: bool = True
    noise_level: float = 0.6
    noise_dof_pos: float = 0.05
    noise_dof_vel: float = 0.5
    noise_ang_vel: float = 0.1
    noise_quat: float = 0.03
    reward_scales: Dict[str, float] = field(default_factory=lambda: dict(XBOT_REWARD_SCALES))

    @property
    def So(self) -> int:
        return 11 + 3 * self.D

    @property
    def Sp(self) -> int:
        return 25 + 4 * self.D

    @property
    def num_obs(self) -> int:
        return self.frame_stack * self.So

    @property
    def num_priv(self) -> int:
        return self.c_frame_stack * self.Sp

    def active_rewards(self) -> List[str]:
        """Alphabetical (class_to_dict walks dir(), utils/helpers.py:43-58), zero scales dropped,
        'termination' skipped (humanoid_env.py:207-221)."""
        return [k for k in sorted(self.reward_scales) if self.reward_scales[k] != 0 and k != "termination"]

    def scaled(self, name: str) -> float:
        """scale * dt in python double, as `_prepare_reward_function` does (:212)."""
        return self.reward_scales[name] * self.dt

    def noise_scale_vec(self) -> torch.Tensor:
        """`_get_noise_scale_vec` (:748-768).  For D=18 the reference's hard-coded slices are
        misaligned with the 65-wide frame; they are reproduced as written."""
        v = torch.zeros(self.So)
        if self.D == 18:
            v[5:23] = self.noise_dof_pos * self.dof_pos_scale
            v[23:41] = self.noise_dof_vel * self.dof_vel_scale
            v[41:53] = 0.0
            v[53:56] = self.noise_ang_vel * self.ang_vel_scale
            v[56:59] = self.noise_quat * self.quat_scale
        else:
            D = self.D
            v[5:5 + D] = self.noise_dof_pos * self.dof_pos_scale
            v[5 + D:5 + 2 * D] = self.noise_dof_vel * self.dof_vel_scale
            v[5 + 3 * D:8 + 3 * D] = self.noise_ang_vel * self.ang_vel_scale
            v[8 + 3 * D:11 + 3 * D] = self.noise_quat * self.quat_scale
        return v


def xbot_spec(D: int = 12) -> EnvSpec:
    if D == 12:
        return EnvSpec()
    if D == 18:  # shipped fork: 6 arm DOFs/bodies before the legs
        return EnvSpec(D=18, leg_offset=6, num_bodies=19, feet=(12, 18), knees=(10, 16))
    raise ValueError(D)


class EnvState:
    """Persistent tensors of the env math (names as in `_init_buffers`, :228-312, and `__init__` :174-175)."""

    def __init__(self, spec: EnvSpec, N: int, default_dof_pos: Optional[torch.Tensor] = None):
        D, z = spec.D, torch.zeros
        self.N = N
        self.actions = z(N, D)
        self.last_actions = z(N, D)
        self.last_last_actions = z(N, D)
        self.last_dof_vel = z(N, D)
        self.last_root_vel = z(N, 6)
        self.feet_air_time = z(N, 2)
        self.last_contacts = z(N, 2, dtype=torch.bool)
        self.feet_height = z(N, 2)
        self.last_feet_z = torch.full((N, 2), 0.05)      # scalar 0.05 in the reference (:174), broadcasts
        self.ref_dof_pos = z(N, D)
        self.episode_length_buf = z(N, dtype=torch.int64)
        self.obs_hist = z(N, spec.frame_stack, spec.So)   # oldest -> newest
        self.crit_hist = z(N, spec.c_frame_stack, spec.Sp)
        self.episode_sums = z(len(spec.active_rewards()), N)
        self.default_dof_pos = z(1, D) if default_dof_pos is None else default_dof_pos.reshape(1, D).clone()

    def clone(self) -> "EnvState":
        import copy
        return copy.deepcopy(self)


def _phase(spec: EnvSpec, ep_len: torch.Tensor) -> torch.Tensor:
    return ep_len * spec.dt / spec.cycle_time                                   # :692


def _stance_mask(spec: EnvSpec, ep_len: torch.Tensor) -> torch.Tensor:
    s = torch.sin(2 * torch.pi * _phase(spec, ep_len))                          # :705
    m = torch.zeros(ep_len.shape[0], 2)
    m[:, 0] = s >= 0
    m[:, 1] = s < 0
    m[torch.abs(s) < 0.1] = 1
    return m


def _ref_dof_pos(spec: EnvSpec, ep_len: torch.Tensor) -> torch.Tensor:
    s = torch.sin(2 * torch.pi * _phase(spec, ep_len))                          # :721-722
    ref = torch.zeros(ep_len.shape[0], spec.D)
    s1, s2, o = spec.target_joint_pos_scale, 2 * spec.target_joint_pos_scale, spec.leg_offset
    sl = s.clone()
    sl[sl > 0] = 0
    ref[:, o + 2] = sl * s1
    ref[:, o + 3] = sl * s2
    ref[:, o + 4] = sl * s1
    sr = s.clone()
    sr[sr < 0] = 0
    ref[:, o + 8] = sr * s1
    ref[:, o + 9] = sr * s2
    ref[:, o + 10] = sr * s1
    ref[torch.abs(s) < 0.1] = 0
    return ref


@dataclass
class CallbackSpec:
    """Config scalars of `_post_physics_step_callback` / `_resample_commands` / `_push_robots` / the `extras` part of
    `reset_idx` (humanoid_config.py:331-337,349-370,64; derived values as humanoid_env.py:188-191,1005)."""
    resampling_steps: int = 800            # int(commands.resampling_time / dt) = int(8 / 0.01)
    heading_command: bool = True
    lin_vel_x: tuple = (-0.3, 0.6)
    lin_vel_y: tuple = (-0.3, 0.3)
    ang_vel_yaw: tuple = (-0.3, 0.3)
    heading: tuple = (-3.14, 3.14)
    push_robots: bool = True
    push_interval: float = 400.0           # np.ceil(push_interval_s / dt) = ceil(4 / 0.01)
    max_push_vel_xy: float = 0.2
    max_push_ang_vel: float = 0.4
    max_episode_length_s: float = 24.0


def rand_float(lower: float, upper: float, u: torch.Tensor) -> torch.Tensor:
    """`torch_rand_float` of Isaac Gym: (upper - lower) * torch.rand(shape) + lower, with the U[0,1) draw `u` given."""
    return (upper - lower) * u + lower


def wrap_to_pi(angles: torch.Tensor) -> torch.Tensor:
    """humanoid/utils/math.py:46-49."""
    angles = angles % (2 * math.pi)
    return angles - 2 * math.pi * (angles > math.pi)


def resample_commands(cb: CallbackSpec, commands: torch.Tensor, ids: torch.Tensor, u: torch.Tensor) -> None:
    """`_resample_commands` (humanoid_env.py:1018-1032) for env ids `ids`; u [N,3] holds each env's three U[0,1) draws
    (x, y, heading-or-yaw).  In place."""
    if len(ids) == 0:
        return
    commands[ids, 0] = rand_float(cb.lin_vel_x[0], cb.lin_vel_x[1], u[ids, 0])
    commands[ids, 1] = rand_float(cb.lin_vel_y[0], cb.lin_vel_y[1], u[ids, 1])
    if cb.heading_command:
        commands[ids, 3] = rand_float(cb.heading[0], cb.heading[1], u[ids, 2])
    else:
        commands[ids, 2] = rand_float(cb.ang_vel_yaw[0], cb.ang_vel_yaw[1], u[ids, 2])
    commands[ids, :2] *= (torch.norm(commands[ids, :2], dim=1) > 0.2).unsqueeze(1)          # :1032


def post_physics_callback(cb: CallbackSpec, ep_len: torch.Tensor, common_step_counter: int, quat: torch.Tensor,
                          commands: torch.Tensor, root_states: torch.Tensor, rand_push_force: torch.Tensor,
                          rand_push_torque: torch.Tensor, u_cmd: torch.Tensor, u_push: torch.Tensor) -> bool:
    """`_post_physics_step_callback` (humanoid_env.py:1001-1016) + `_push_robots` (:665-681), in place on commands /
    root_states / rand_push_force / rand_push_torque.  ep_len and common_step_counter are the already incremented values
    (:780-781).  u_cmd [N,3], u_push [N,5] are U[0,1) draws.  Returns whether this step pushed."""
    ids = (ep_len % cb.resampling_steps == 0).nonzero(as_tuple=False).flatten()
    resample_commands(cb, commands, ids, u_cmd)
    if cb.heading_command:
        forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat(quat.shape[0], 1)
        forward = quat_apply(quat, forward_vec)
        heading = torch.atan2(forward[:, 1], forward[:, 0])
        commands[:, 2] = torch.clip(0.5 * wrap_to_pi(commands[:, 3] - heading), -1.0, 1.0)
    pushed = cb.push_robots and (common_step_counter % cb.push_interval == 0)
    if pushed:
        rand_push_force[:, :2] = rand_float(-cb.max_push_vel_xy, cb.max_push_vel_xy, u_push[:, :2])
        root_states[:, 7:9] = rand_push_force[:, :2]
        rand_push_torque[:] = rand_float(-cb.max_push_ang_vel, cb.max_push_ang_vel, u_push[:, 2:5])
        root_states[:, 10:13] = rand_push_torque
    return bool(pushed)


def episode_means(cb: CallbackSpec, episode_sums: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """extras["episode"]["rew_<name>"] of reset_idx (:1141-1143): mean over the reset envs of every episode sum,
    divided by max_episode_length_s.  episode_sums [K,N] -> [K]."""
    return torch.stack([torch.mean(episode_sums[k][ids]) / cb.max_episode_length_s for k in range(episode_sums.shape[0])])


def env_step(spec: EnvSpec, st: EnvState, sim: Dict[str, torch.Tensor], noise: Optional[torch.Tensor],
             noise_vec: Optional[torch.Tensor] = None, callback: Optional[Dict] = None):
    """One post-physics step.  `sim` holds the Isaac-Gym-shaped tensors for this step:
    root_states [N,13], dof_pos [N,D], dof_vel [N,D], contact_forces [N,B,3], rigid_state [N,B,13],
    commands [N,4], actions [N,D] (already delayed/noised/clipped, :624-635), torques [N,D],
    rand_push_force [N,3], rand_push_torque [N,3], env_frictions [N,1], body_mass [N,1].
    `noise` [N,So] is the N(0,1) draw of `randn_like` (:877) or None.
    Returns dict(obs, priv, rew, reset, time_out, terms[K,N]).  Mutates `st`.
    `callback` = dict(spec=CallbackSpec, common_step_counter=int (already incremented), u_cmd [N,3],
    u_push [N,5], u_cmd_reset [N,3]): additionally runs `_post_physics_step_callback` between the base-frame quantities
    and check_termination, and the command resampling + episode means of `reset_idx`; commands, rand_push_force and
    rand_push_torque then live in `st` (sim's are ignored) and the returned dict gains commands, root_states,
    pushed, episode_means (None when no env reset: reset_idx returns early, :1113-1114)."""
    N = st.N
    feet, knees = list(spec.feet), list(spec.knees)
    if noise_vec is None:
        noise_vec = spec.noise_scale_vec()
    root, dof_pos, dof_vel = sim["root_states"], sim["dof_pos"], sim["dof_vel"]
    cf, rigid, cmd = sim["contact_forces"], sim["rigid_state"], sim["commands"]
    st.actions = sim["actions"].clone()
    torques = sim["torques"]

    st.episode_length_buf += 1                                                  # :780
    quat = root[:, 3:7]
    base_lin_vel = quat_rotate_inverse(quat, root[:, 7:10])                     # :785
    base_ang_vel = quat_rotate_inverse(quat, root[:, 10:13])                    # :786
    gravity = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
    proj_grav = quat_rotate_inverse(quat, gravity)                              # :787
    euler = euler_xyz_wrapped(quat)                                             # :788
    push_force, push_torque, pushed, ep_means = sim["rand_push_force"], sim["rand_push_torque"], False, None
    if callback is not None:                                                    # :791 (base velocities above are PRE-push)
        if not hasattr(st, "commands"):
            st.commands, st.rand_push_force, st.rand_push_torque = torch.zeros(N, 4), torch.zeros(N, 3), torch.zeros(N, 3)
        root = root.clone()
        cmd, push_force, push_torque = st.commands, st.rand_push_force, st.rand_push_torque
        pushed = post_physics_callback(callback["spec"], st.episode_length_buf, callback["common_step_counter"], quat, cmd,
                                       root, push_force, push_torque, callback["u_cmd"], callback["u_push"])

    # check_termination :811-816
    reset = torch.any(torch.norm(cf[:, list(spec.termination_bodies), :], dim=-1) > 1.0, dim=1)
    time_out = st.episode_length_buf > spec.max_episode_length
    reset = reset | time_out

    # ---- rewards :1170-1437 (stance uses the incremented, not yet reset, episode length)
    stance = _stance_mask(spec, st.episode_length_buf)
    contact = cf[:, feet, 2] > 5.0
    o = spec.leg_offset
    r: Dict[str, torch.Tensor] = {}
    a, la, lla = st.actions, st.last_actions, st.last_last_actions
    r["action_smoothness"] = (torch.sum(torch.square(la - a), dim=1)
                              + torch.sum(torch.square(a + lla - 2 * la), dim=1)
                              + 0.05 * torch.sum(torch.abs(a), dim=1))
    r["base_acc"] = torch.exp(-torch.norm(st.last_root_vel - root[:, 7:13], dim=1) * 3)
    meas = torch.sum(rigid[:, feet, 2] * stance, dim=1) / torch.sum(stance, dim=1)
    r["base_height"] = torch.exp(-torch.abs(root[:, 2] - (meas - 0.05) - spec.base_height_target) * 100)
    r["collision"] = torch.sum(1.0 * (torch.norm(cf[:, list(spec.penalised_bodies), :], dim=-1) > 0.1), dim=1)
    jd = dof_pos - st.default_dof_pos
    yr = torch.norm(jd[:, o:o + 2], dim=1) + torch.norm(jd[:, o + 6:o + 8], dim=1)
    yr = torch.clamp(yr - 0.1, 0, 50)
    r["default_joint_pos"] = torch.exp(-yr * 100) - 0.01 * torch.norm(jd, dim=1)
    r["dof_acc"] = torch.sum(torch.square((st.last_dof_vel - dof_vel) / spec.dt), dim=1)
    r["dof_vel"] = torch.sum(torch.square(dof_vel), dim=1)
    # feet_air_time :1216-1230 (stateful)
    contact_filt = torch.logical_or(torch.logical_or(contact, stance), st.last_contacts)
    st.last_contacts = contact
    first_contact = (st.feet_air_time > 0.0) * contact_filt
    st.feet_air_time += spec.dt
    air = st.feet_air_time.clamp(0, 0.5) * first_contact
    st.feet_air_time *= ~contact_filt
    r["feet_air_time"] = air.sum(dim=1)
    # feet_clearance :1342-1363 (stateful)
    feet_z = rigid[:, feet, 2] - 0.05
    st.feet_height += feet_z - st.last_feet_z
    st.last_feet_z = feet_z
    swing = 1 - stance
    rew_pos = torch.abs(st.feet_height - spec.target_feet_height) < 0.01
    r["feet_clearance"] = torch.sum(rew_pos * swing, dim=1)
    st.feet_height *= ~contact
    r["feet_contact_forces"] = torch.sum(
        (torch.norm(cf[:, feet, :], dim=-1) - spec.max_contact_force).clip(0, 400), dim=1)
    r["feet_contact_number"] = torch.mean(torch.where(contact == stance, 1, -0.3), dim=1)

    def _dist_reward(idx, max_d):
        pos = rigid[:, idx, :2]
        d = torch.norm(pos[:, 0, :] - pos[:, 1, :], dim=1)
        d_min = torch.clamp(d - spec.min_dist, -0.5, 0.0)
        d_max = torch.clamp(d - max_d, 0, 0.5)
        return (torch.exp(-torch.abs(d_min) * 100) + torch.exp(-torch.abs(d_max) * 100)) / 2

    r["feet_distance"] = _dist_reward(feet, spec.max_dist)
    slip = torch.sqrt(torch.norm(rigid[:, feet, 10:12], dim=2))
    slip *= contact
    r["foot_slip"] = torch.sum(slip, dim=1)
    jn = torch.norm(dof_pos - st.ref_dof_pos, dim=1)                           # stale ref (:1176)
    r["joint_pos"] = torch.exp(-2 * jn) - 0.2 * jn.clamp(0, 0.5)
    r["knee_distance"] = _dist_reward(knees, spec.max_dist / 2)
    vx, cx = base_lin_vel[:, 0], cmd[:, 0]
    too_low = torch.abs(vx) < 0.5 * torch.abs(cx)
    too_high = torch.abs(vx) > 1.2 * torch.abs(cx)
    ls = torch.zeros_like(vx)
    ls[too_low] = -1.0
    ls[too_high] = 0.0
    ls[~(too_low | too_high)] = 1.2
    ls[torch.sign(vx) != torch.sign(cx)] = -2.0
    r["low_speed"] = ls * (cx.abs() > 0.1)
    r["orientation"] = (torch.exp(-torch.sum(torch.abs(euler[:, :2]), dim=1) * 10)
                        + torch.exp(-torch.norm(proj_grav[:, :2], dim=1) * 20)) / 2.0
    r["torques"] = torch.sum(torch.square(torques), dim=1)
    e_l = torch.norm(cmd[:, :2] - base_lin_vel[:, :2], dim=1)
    e_a = torch.abs(cmd[:, 2] - base_ang_vel[:, 2])
    r["track_vel_hard"] = (torch.exp(-e_l * 10) + torch.exp(-e_a * 10)) / 2.0 - 0.2 * (e_l + e_a)
    r["tracking_ang_vel"] = torch.exp(-torch.square(cmd[:, 2] - base_ang_vel[:, 2]) * spec.tracking_sigma)
    r["tracking_lin_vel"] = torch.exp(
        -torch.sum(torch.square(cmd[:, :2] - base_lin_vel[:, :2]), dim=1) * spec.tracking_sigma)
    r["vel_mismatch_exp"] = (torch.exp(-torch.square(base_lin_vel[:, 2]) * 10)
                             + torch.exp(-torch.norm(base_ang_vel[:, :2], dim=1) * 5.0)) / 2.0

    names = spec.active_rewards()
    rew = torch.zeros(N)
    terms = torch.zeros(len(names), N)
    for k, name in enumerate(names):                                            # :894-898
        term = r[name] * spec.scaled(name)
        rew += term
        st.episode_sums[k] += term
        terms[k] = term
    if spec.only_positive_rewards:
        rew = torch.clip(rew, min=0.0)

    # ---- reset_idx buffer part :1132-1139, :1141-1144, :1160-1163
    ids = reset.nonzero(as_tuple=False).flatten()
    if len(ids) > 0:
        if callback is not None:
            resample_commands(callback["spec"], cmd, ids, callback["u_cmd_reset"])     # :1127
            ep_means = episode_means(callback["spec"], st.episode_sums, ids)           # :1141-1143 (before the zeroing)
        st.last_last_actions[ids] = 0.0
        st.actions[ids] = 0.0
        st.last_actions[ids] = 0.0
        st.last_dof_vel[ids] = 0.0
        st.feet_air_time[ids] = 0.0
        st.episode_length_buf[ids] = 0
        st.episode_sums[:, ids] = 0.0
        st.obs_hist[ids] *= 0
        st.crit_hist[ids] *= 0

    # ---- compute_observations :818-887
    phase = _phase(spec, st.episode_length_buf)
    st.ref_dof_pos = _ref_dof_pos(spec, st.episode_length_buf)
    sin_p = torch.sin(2 * torch.pi * phase).unsqueeze(1)
    cos_p = torch.cos(2 * torch.pi * phase).unsqueeze(1)
    stance = _stance_mask(spec, st.episode_length_buf)
    cmd_scale = torch.tensor([spec.lin_vel_scale, spec.lin_vel_scale, spec.ang_vel_scale])
    command_input = torch.cat((sin_p, cos_p, cmd[:, :3] * cmd_scale), dim=1)
    q = (dof_pos - st.default_dof_pos) * spec.dof_pos_scale
    dq = dof_vel * spec.dof_vel_scale
    diff = dof_pos - st.ref_dof_pos
    priv = torch.cat((command_input, q, dq, st.actions, diff,
                      base_lin_vel * spec.lin_vel_scale, base_ang_vel * spec.ang_vel_scale,
                      euler * spec.quat_scale, push_force[:, :2], push_torque,
                      sim["env_frictions"], sim["body_mass"] / 30.0, stance, contact), dim=-1)
    frame = torch.cat((command_input, q, dq, st.actions, base_ang_vel * spec.ang_vel_scale,
                       euler * spec.quat_scale), dim=-1)
    if spec.add_noise and noise is not None:
        frame = frame.clone() + noise * noise_vec * spec.noise_level           # :877
    st.obs_hist = torch.cat((st.obs_hist[:, 1:], frame.unsqueeze(1)), dim=1)   # deque(maxlen).append
    st.crit_hist = torch.cat((st.crit_hist[:, 1:], priv.unsqueeze(1)), dim=1)
    obs_buf = st.obs_hist.reshape(N, -1)
    priv_buf = st.crit_hist.reshape(N, -1)

    # ---- hand-over :802-806
    st.last_last_actions = st.last_actions.clone()
    st.last_actions = st.actions.clone()
    st.last_dof_vel = dof_vel.clone()
    st.last_root_vel = root[:, 7:13].clone()

    obs_buf = torch.clip(obs_buf, -spec.clip_obs, spec.clip_obs)               # :654-657
    priv_buf = torch.clip(priv_buf, -spec.clip_obs, spec.clip_obs)
    out = dict(obs=obs_buf, priv=priv_buf, rew=rew, reset=reset, time_out=time_out, terms=terms,
               base_lin_vel=base_lin_vel, base_ang_vel=base_ang_vel, euler=euler, proj_grav=proj_grav)
    if callback is not None:
        out.update(commands=cmd.clone(), root_states=root.clone(), pushed=pushed, episode_means=ep_means,
                   rand_push_force=push_force.clone(), rand_push_torque=push_torque.clone())
    return out


def synthetic_sim_state(spec: EnvSpec, N: int, gen: torch.Generator, upright: bool = True,
                        base_contact_frac: float = 0.02) -> Dict[str, torch.Tensor]:
    """Isaac-Gym-shaped synthetic tensors for one step."""
    D, B = spec.D, spec.num_bodies
    rn = lambda *s: torch.randn(*s, generator=gen)
    root = torch.zeros(N, 13)
    root[:, :3] = torch.tensor([0.0, 0.0, 0.95]) + 0.1 * rn(N, 3)
    if upright:
        quat = torch.cat((0.1 * rn(N, 3), torch.ones(N, 1)), dim=1)
    else:
        quat = rn(N, 4)
    root[:, 3:7] = quat / quat.norm(dim=1, keepdim=True)
    root[:, 7:13] = 0.3 * rn(N, 6)
    dof_state = 0.2 * rn(N * D, 2)
    cf = (20.0 * rn(N, B, 3)).abs()
    base_hit = torch.rand(N, generator=gen) < base_contact_frac
    cf[:, list(spec.termination_bodies), :] = torch.where(
        base_hit[:, None, None], cf[:, list(spec.termination_bodies), :], 0.05 * rn(N, len(spec.termination_bodies), 3))
    rigid = 0.3 * rn(N, B, 13)
    rigid[:, list(spec.feet), 2] = 0.15 * torch.rand(N, 2, generator=gen)
    return dict(
        root_states=root, dof_state=dof_state,
        dof_pos=dof_state.view(N, D, 2)[..., 0], dof_vel=dof_state.view(N, D, 2)[..., 1],
        contact_forces=cf, rigid_state=rigid, commands=0.3 * rn(N, 4),
        actions=rn(N, D), torques=10.0 * rn(N, D),
        rand_push_force=0.2 * rn(N, 3), rand_push_torque=0.4 * rn(N, 3),
        env_frictions=0.1 + 1.9 * torch.rand(N, 1, generator=gen),
        body_mass=20.0 + 10.0 * torch.rand(N, 1, generator=gen))


# ----------------------------------------------------------------------------- pre-physics math
def mix_actions(raw_actions: torch.Tensor, prev_actions: torch.Tensor, delay: torch.Tensor, noise: torch.Tensor,
                dyn_rand: float = 0.02, clip_actions: float = 18.0) -> torch.Tensor:
    """`XBotLFreeEnv.step` :624-636: action delay mix with delay ~ U[0,1) [N,1], multiplicative noise with
    noise ~ N(0,1) [N,D] (`dynamic_randomization`, humanoid_config.py domain_rand), clip.  Returns the new self.actions."""
    actions = (1 - delay) * raw_actions + delay * prev_actions
    actions = actions + dyn_rand * noise * actions
    return torch.clip(actions, -clip_actions, clip_actions)


def compute_torques(actions: torch.Tensor, dof_pos: torch.Tensor, dof_vel: torch.Tensor, p_gains: torch.Tensor,
                    d_gains: torch.Tensor, default_dof_pos: torch.Tensor, torque_limits: torch.Tensor,
                    action_scale: float = 0.25) -> torch.Tensor:
    """`_compute_torques` :910-925 (PD law + clip)."""
    actions_scaled = actions * action_scale
    torques = p_gains * (actions_scaled + default_dof_pos - dof_pos) - d_gains * dof_vel
    return torch.clip(torques, -torque_limits, torque_limits)
