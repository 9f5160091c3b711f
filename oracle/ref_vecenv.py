"""`VecEnv`-shaped (reference humanoid/algo/vec_env.py:39-63) wrapper around the reference's own env tensor math
(`ref_harness.RefEnv`: the unmodified text of humanoid_env.py bound to synthetic Isaac-Gym-shaped tensors), so that the
UNMODIFIED `OnPolicyRunner` can be driven end to end without Isaac Gym.  TEST / BASELINE INFRASTRUCTURE ONLY.

It also carries the two attributes `OnPolicyRunner.learn` reads but `XBotLFreeEnv` never defines
(`env.course_gain`, `env.course_ratio`, on_policy_runner.py:160-162; only `LeggedRobot` has them, legged_robot.py:84-85).
"""
from __future__ import annotations

import torch

from . import env_oracle as eo
from . import ref_harness as rh


def train_cfg_dict(num_steps_per_env=None, save_interval=None, max_iterations=None):
    """`class_to_dict(XBotLCfgPPO())` exactly as task_registry.make_alg_runner builds it (task_registry.py:140-154)."""
    rh._install_stubs()
    from humanoid.envs.custom.humanoid_config import XBotLCfgPPO
    from humanoid.utils.helpers import class_to_dict
    d = class_to_dict(XBotLCfgPPO())
    if num_steps_per_env is not None:
        d["runner"]["num_steps_per_env"] = num_steps_per_env
    if save_interval is not None:
        d["runner"]["save_interval"] = save_interval
    if max_iterations is not None:
        d["runner"]["max_iterations"] = max_iterations
    return d


class RefSyntheticVecEnv:
    def __init__(self, N: int, D: int = 12, device="cpu", pool: int = 4, seed: int = 0):
        self.spec = eo.xbot_spec(D)
        self.device = device
        self.ref = rh.RefEnv(D, N, self.spec, device=device)
        self.num_envs, self.num_actions = N, D
        self.num_obs, self.num_privileged_obs = self.spec.num_obs, self.spec.num_priv
        self.max_episode_length = self.ref.e.max_episode_length
        self.course_gain, self.course_ratio = 1.0, 1.0
        with torch.device(device):
            g = torch.Generator(device=device).manual_seed(seed)
            self.pool = [eo.synthetic_sim_state(self.spec, N, g) for _ in range(pool)]
            out = self.ref.step(self.pool[0], torch.randn(N, self.spec.So, generator=g))
        self.obs_buf, self.privileged_obs_buf = out["obs"], out["priv"]
        self.extras = {}
        self._i = 0

    @property
    def episode_length_buf(self):
        return self.ref.e.episode_length_buf

    @episode_length_buf.setter
    def episode_length_buf(self, v):
        self.ref.e.episode_length_buf = v

    def get_observations(self):
        return self.obs_buf

    def get_privileged_observations(self):
        return self.privileged_obs_buf

    def reset(self):
        return self.obs_buf, self.privileged_obs_buf

    def step(self, actions):
        self._i += 1
        sim = dict(self.pool[self._i % len(self.pool)])
        sim["actions"] = torch.clip(actions, -18.0, 18.0)                     # step() :633-635
        with torch.device(self.device):
            out = self.ref.step(sim, torch.randn(self.num_envs, self.spec.So))
        self.obs_buf, self.privileged_obs_buf = out["obs"], out["priv"]
        self.extras["time_outs"] = out["time_out"]
        return self.obs_buf, self.privileged_obs_buf, out["rew"], out["reset"], self.extras
