"""Small, complete pass over every kernel of the hot path for compute-sanitizer (tools/sanitize.sh)."""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from humanoid_gym_with_comments_b200 import _lib
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

dev = "cuda:0"
N, T = 256, 4
torch.manual_seed(0)
lib = _lib.load()
l0 = lib.hg_launch_count()
for ring in (True, False):
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
    alg = PPO(ac, num_learning_epochs=1, num_mini_batches=2, clip_param=0.2, gamma=0.994, lam=0.9, entropy_coef=0.001,
              learning_rate=1e-5, schedule="adaptive", desired_kl=0.01, device=dev)
    alg.init_storage(N, T, [705], [219], [12])
    alg.torch_rng = False
    env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), dev, seed=1, torch_rng_noise=False, callbacks=not ring)
    env.bind_storage(alg.storage, gamma=alg.gamma, ring=ring)
    obs, cobs = env.get_observations(), env.get_privileged_observations()
    for it in range(2):
        with torch.inference_mode():
            for t in range(T):
                a = alg.act(obs, cobs)
                obs, cobs, rew, dones, infos = env.step(a)
                alg.process_env_step(rew, dones, infos)
            alg.compute_returns(cobs)
        o = alg.storage.observations          # ring: re-stacked from the archive
        losses = alg.update()
    torch.cuda.synchronize()
    print("workload", "ring" if ring else "stacked+callbacks", "losses", [round(float(x), 5) for x in losses])
# generic (unbound) path: stacked env buffers -> ingest
env.bind_storage(None)
a = alg.act(env.get_observations(), env.get_privileged_observations())
obs, cobs, rew, dones, infos = env.step(a)
alg.process_env_step(rew, dones, infos)
torch.cuda.synchronize()
print("workload done:", lib.hg_launch_count() - l0, "launches")
