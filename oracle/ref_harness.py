"""Loads the UNMODIFIED reference -- from /root/reference in the build container, from the byte-identical copy
`oracle/_ref/` (recipe: oracle/build_ref.py, git-ignored, travels with the snapshot) on the GPU box -- so that
`make_golden.py` can pin the oracle and `bench.py --impl reference` / the runner test can execute the real thing.
TEST / BASELINE INFRASTRUCTURE ONLY.

* learner: `humanoid.algo.ppo` imports cleanly (needs wandb + tensorboard, both present).
* env math: `humanoid_env.py` imports `isaacgym` (closed, absent) and broken package __init__s,
  so a stub `isaacgym` exporting the restated torch_utils helpers and bare
  package modules are installed in `sys.modules` first; the unbound methods of `XBotLFreeEnv`
  are then bound onto a namespace that carries synthetic Isaac-Gym-shaped tensors.
* 12-DOF layout: the reference text with the 13 single-occurrence `+6` replacements, exec'd.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from collections import deque

import numpy as np
import torch

def _find_ref_root() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("HG_REFERENCE_ROOT"), "/root/reference", os.path.join(here, "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "humanoid", "algo", "ppo")):
            return c
    return "/root/reference"


REF_ROOT = _find_ref_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "humanoid", "algo", "ppo"))


def load_runner():
    """-> the reference's `humanoid.algo.ppo.on_policy_runner` module (imports wandb + tensorboard)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        import importlib       # (`from .ppo import *` in humanoid/algo/__init__.py shadows the package attribute)
        opr = importlib.import_module("humanoid.algo.ppo.on_policy_runner")
    return opr


def load_learner():
    """-> (PPO, ActorCritic, RolloutStorage) of the reference."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with contextlib.redirect_stdout(io.StringIO()):
        from humanoid.algo.ppo import PPO, ActorCritic, RolloutStorage  # noqa
    return PPO, ActorCritic, RolloutStorage


def _install_stubs():
    from . import isaac_math
    if "isaacgym" in sys.modules and getattr(sys.modules["isaacgym"], "_hg_stub", False):
        return
    pkg = types.ModuleType("isaacgym")
    pkg.__path__ = []
    pkg._hg_stub = True
    for sub in ("gymapi", "gymtorch", "gymutil", "terrain_utils"):
        m = types.ModuleType("isaacgym." + sub)
        if sub == "gymtorch":
            m.unwrap_tensor = lambda t: t          # _push_robots hands root_states back to the simulator (:680-681)
        setattr(pkg, sub, m)
        sys.modules["isaacgym." + sub] = m
    tu = types.ModuleType("isaacgym.torch_utils")
    tu.torch, tu.np = torch, np
    tu.quat_rotate_inverse = isaac_math.quat_rotate_inverse
    tu.get_euler_xyz = isaac_math.get_euler_xyz
    tu.to_torch = isaac_math.to_torch
    tu.get_axis_params = isaac_math.get_axis_params
    tu.quat_apply = isaac_math.quat_apply
    tu.normalize = lambda x, eps=1e-9: x / x.norm(p=2, dim=-1).clamp(min=eps).unsqueeze(-1)
    tu.torch_rand_float = lambda lo, hi, shape, device: (hi - lo) * torch.rand(*shape, device=device) + lo
    tu.__all__ = ["torch", "np", "quat_rotate_inverse", "get_euler_xyz", "to_torch", "get_axis_params",
                  "quat_apply", "normalize", "torch_rand_float"]
    pkg.torch_utils = tu
    sys.modules["isaacgym"] = pkg
    sys.modules["isaacgym.torch_utils"] = tu
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import humanoid  # real top-level package (only defines LEGGED_GYM_ROOT_DIR)
    for name in ("humanoid.envs", "humanoid.envs.base", "humanoid.envs.custom", "humanoid.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
            sys.modules[name] = m


_TWELVE_DOF_EDITS = [
    ("[:, 2+6]", "[:, 2]"), ("[:, 3+6]", "[:, 3]"), ("[:, 4+6]", "[:, 4]"),
    ("[:, 8+6]", "[:, 8]"), ("[:, 9+6]", "[:, 9]"), ("[:, 10+6]", "[:, 10]"),
    ("noise_vec[5: 17+6]", "noise_vec[5: 17]"), ("noise_vec[17+6: 29+6*2]", "noise_vec[17: 29]"),
    ("noise_vec[29+6*2: 41+6*2]", "noise_vec[29: 41]"), ("noise_vec[41+6*2: 44+6*2]", "noise_vec[41: 44]"),
    ("noise_vec[44+6*2: 47+6*2]", "noise_vec[44: 47]"),
    ("joint_diff[:, 6:8]  #", "joint_diff[:, :2]  #"), ("joint_diff[:, 12: 14]  #", "joint_diff[:, 6: 8]  #"),
]


def load_env_class(D: int):
    """-> (XBotLFreeEnv class, XBotLCfg class) for the shipped layout (D=18) or the 12-DOF text."""
    _install_stubs()
    with contextlib.redirect_stdout(io.StringIO()):
        import humanoid.envs.custom.humanoid_env as shipped
    if D == 18:
        return shipped.XBotLFreeEnv, shipped.XBotLCfg
    path = os.path.join(REF_ROOT, "humanoid", "envs", "custom", "humanoid_env.py")
    text = open(path, encoding="utf-8").read()
    for old, new in _TWELVE_DOF_EDITS:
        assert text.count(old) == 1, (old, text.count(old))
        text = text.replace(old, new)
    mod = types.ModuleType("humanoid.envs.custom.humanoid_env_12dof")
    mod.__package__ = "humanoid.envs.custom"
    mod.__file__ = path
    exec(compile(text, path, "exec"), mod.__dict__)
    return mod.XBotLFreeEnv, mod.XBotLCfg


_BOUND = ["_get_phase", "_get_gait_phase", "compute_ref_state", "_get_noise_scale_vec", "check_termination",
          "compute_observations", "compute_reward", "_prepare_reward_function", "_post_physics_step_callback",
          "_resample_commands", "_push_robots"]


class _Replay:
    """Feeds `torch_rand_float` the given U[0,1) draws: the reference draws `len(env_ids)` numbers per call, the
    kernels take one draw per env, so each queued call carries the rows it must return."""

    def __init__(self):
        self.queue = []

    def __call__(self, lower, upper, shape, device):
        u = self.queue.pop(0)
        assert tuple(u.shape) == tuple(shape), (u.shape, shape)
        return (upper - lower) * u + lower


class RefEnv:
    """The reference env math bound to synthetic tensors (no Isaac Gym).  `step(sim, noise)` runs
    exactly the tensor part of `post_physics_step` (:780-806) + the clips of `step` (:654-657)."""

    def __init__(self, D: int, N: int, spec, device="cpu"):
        """device="cuda:0": the same unmodified text on torch-CUDA (every factory call below and in the reference's
        methods lands on `device` through torch's default-device context / `self.device`)."""
        self.device = device
        with torch.device(device):
            self._init(D, N, spec, device)

    def _init(self, D, N, spec, device):
        cls, Cfg = load_env_class(D)
        self.cls = cls
        e = types.SimpleNamespace()
        cfg = Cfg()
        cfg.env.num_actions = D
        cfg.env.num_single_obs = spec.So
        cfg.env.single_num_privileged_obs = spec.Sp
        e.cfg, e.num_envs, e.device, e.dt, e.num_dof, e.num_actions = cfg, N, device, spec.dt, D, D
        e.obs_scales = cfg.normalization.obs_scales
        e.max_episode_length = np.ceil(cfg.env.episode_length_s / e.dt)
        from humanoid.utils.helpers import class_to_dict
        e.reward_scales = class_to_dict(cfg.rewards.scales)
        for name in dir(cls):
            if name.startswith("_reward_") or name in _BOUND:
                setattr(e, name, types.MethodType(getattr(cls, name), e))
        e.feet_indices = torch.tensor(spec.feet)
        e.knee_indices = torch.tensor(spec.knees)
        e.penalised_contact_indices = torch.tensor(spec.penalised_bodies)
        e.termination_contact_indices = torch.tensor(spec.termination_bodies)
        e.commands_scale = torch.tensor([e.obs_scales.lin_vel, e.obs_scales.lin_vel, e.obs_scales.ang_vel])
        e.default_dof_pos = torch.zeros(1, D)
        e.default_joint_pd_target = e.default_dof_pos.clone()
        z = torch.zeros
        e.actions, e.last_actions, e.last_last_actions = z(N, D), z(N, D), z(N, D)
        e.last_dof_vel, e.last_root_vel = z(N, D), z(N, 6)
        e.feet_air_time = z(N, 2)
        e.last_contacts = z(N, 2, dtype=torch.bool)
        e.feet_height = z(N, 2)
        e.last_feet_z = 0.05
        e.gravity_vec = torch.tensor([0.0, 0.0, -1.0]).repeat(N, 1)
        e.episode_length_buf = z(N, dtype=torch.long)
        e.rew_buf = z(N)
        e.obs_history = deque(maxlen=cfg.env.frame_stack)
        e.critic_history = deque(maxlen=cfg.env.c_frame_stack)
        for _ in range(cfg.env.frame_stack):
            e.obs_history.append(z(N, spec.So))
        for _ in range(cfg.env.c_frame_stack):
            e.critic_history.append(z(N, spec.Sp))
        e.noise_scale_vec = e._get_noise_scale_vec(cfg)
        e._prepare_reward_function()
        e.ref_dof_pos = z(N, D)       # the reference first sets it in compute_observations() at init (:177)
        # _post_physics_step_callback / _resample_commands / _push_robots state (:188-191,257-271)
        e.max_episode_length_s = cfg.env.episode_length_s
        cfg.domain_rand.push_interval = np.ceil(cfg.domain_rand.push_interval_s / e.dt)
        e.command_ranges = class_to_dict(cfg.commands.ranges)
        e.commands = z(N, cfg.commands.num_commands)
        e.forward_vec = torch.tensor([1.0, 0.0, 0.0]).repeat(N, 1)
        e.common_step_counter = 0
        e.gym = types.SimpleNamespace(set_actor_root_state_tensor=lambda sim, t: None)
        e.sim = None
        self.replay = _Replay()
        self.e = e
        self._randn = None

    def step(self, sim, noise, callback=None):
        with torch.device(self.device):
            return self._step(sim, noise, callback)

    def _step(self, sim, noise, callback=None):
        """callback = dict(u_cmd [N,3], u_push [N,5], u_cmd_reset [N,3]): also run the reference's
        `_post_physics_step_callback` (with `_resample_commands`, `_push_robots`) and the command resampling + episode
        means of `reset_idx`, replaying the given uniform draws; commands / push buffers then persist in the env."""
        e = self.e
        g = self.cls.__init__.__globals__
        g["torch_rand_float"] = self.replay
        from .isaac_math import quat_rotate_inverse
        get_euler_xyz_tensor = self.cls.__init__.__globals__["get_euler_xyz_tensor"]
        e.root_states, e.dof_pos, e.dof_vel = sim["root_states"], sim["dof_pos"], sim["dof_vel"]
        e.contact_forces, e.rigid_state = sim["contact_forces"], sim["rigid_state"]
        e.actions = sim["actions"].clone()
        e.torques = sim["torques"]
        if callback is None:
            e.commands = sim["commands"]
            e.rand_push_force, e.rand_push_torque = sim["rand_push_force"], sim["rand_push_torque"]
        else:
            e.root_states = e.root_states.clone()
            if not hasattr(e, "rand_push_force"):
                e.rand_push_force, e.rand_push_torque = torch.zeros(e.num_envs, 3), torch.zeros(e.num_envs, 3)
        e.env_frictions, e.body_mass = sim["env_frictions"], sim["body_mass"]
        e.episode_length_buf += 1
        e.base_quat = e.root_states[:, 3:7]
        e.base_lin_vel = quat_rotate_inverse(e.base_quat, e.root_states[:, 7:10])
        e.base_ang_vel = quat_rotate_inverse(e.base_quat, e.root_states[:, 10:13])
        e.projected_gravity = quat_rotate_inverse(e.base_quat, e.gravity_vec)
        e.base_euler_xyz = get_euler_xyz_tensor(e.base_quat)
        ep_means = None
        if callback is not None:
            e.common_step_counter += 1
            ids = (e.episode_length_buf % int(e.cfg.commands.resampling_time / e.dt) == 0).nonzero(as_tuple=False).flatten()
            u = callback["u_cmd"]
            self.replay.queue = [u[ids, 0:1], u[ids, 1:2], u[ids, 2:3]]
            if e.cfg.domain_rand.push_robots and e.common_step_counter % e.cfg.domain_rand.push_interval == 0:
                self.replay.queue += [callback["u_push"][:, :2], callback["u_push"][:, 2:5]]
            e._post_physics_step_callback()
            assert not self.replay.queue
        e.check_termination()
        e.compute_reward()
        ids = e.reset_buf.nonzero(as_tuple=False).flatten()
        if len(ids) > 0 and callback is not None:          # reset_idx :1127 and :1141-1143, the reference's own lines
            u = callback["u_cmd_reset"]
            self.replay.queue = [u[ids, 0:1], u[ids, 1:2], u[ids, 2:3]]
            e._resample_commands(ids)
            ep_means = torch.stack([torch.mean(e.episode_sums[k][ids]) / e.max_episode_length_s for k in e.episode_sums])
        if len(ids) > 0:                                   # tensor part of reset_idx (:1132-1163)
            e.last_last_actions[ids] = 0.0
            e.actions[ids] = 0.0
            e.last_actions[ids] = 0.0
            e.last_dof_vel[ids] = 0.0
            e.feet_air_time[ids] = 0.0
            e.episode_length_buf[ids] = 0
            for k in e.episode_sums:
                e.episode_sums[k][ids] = 0.0
            for i in range(e.obs_history.maxlen):
                e.obs_history[i][ids] *= 0
            for i in range(e.critic_history.maxlen):
                e.critic_history[i][ids] *= 0
        # compute_observations draws randn_like internally (:877): replay the supplied noise
        real_randn_like = torch.randn_like
        torch.randn_like = lambda t, *a, **k: noise if noise is not None else torch.zeros_like(t)
        try:
            e.add_noise = noise is not None
            e.compute_observations()
        finally:
            torch.randn_like = real_randn_like
        e.last_last_actions[:] = torch.clone(e.last_actions[:])
        e.last_actions[:] = e.actions[:]
        e.last_dof_vel[:] = e.dof_vel[:]
        e.last_root_vel[:] = e.root_states[:, 7:13]
        clip = e.cfg.normalization.clip_observations
        obs = torch.clip(e.obs_buf, -clip, clip)
        priv = torch.clip(e.privileged_obs_buf, -clip, clip)
        out = dict(obs=obs, priv=priv, rew=e.rew_buf.clone(), reset=e.reset_buf.clone(),
                   time_out=e.time_out_buf.clone())
        if callback is not None:
            out.update(commands=e.commands.clone(), root_states=e.root_states.clone(), episode_means=ep_means,
                       rand_push_force=e.rand_push_force.clone(), rand_push_torque=e.rand_push_torque.clone())
        return out
