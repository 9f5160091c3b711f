"""B200-native (sm_100a) implementation of humanoid-gym's data-parallel PPO hot path.

Drop-in surface (mirrors `humanoid.algo.ppo` of Rengar-Yang/humanoid-gym-with-comments):
    from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic, RolloutStorage
Env-side fused tensor math: `humanoid_gym_with_comments_b200.envs`.
All arithmetic runs in hand-written CUDA kernels behind the C ABI of include/hgb200.h.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
