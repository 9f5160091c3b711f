"""Drop-in for `humanoid.algo.ppo` of the reference (humanoid/algo/ppo/__init__.py): same three class names."""
from .actor_critic import ActorCritic
from .ppo import PPO
from .rollout_storage import RolloutStorage

__all__ = ["PPO", "ActorCritic", "RolloutStorage"]
