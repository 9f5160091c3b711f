"""One PPO iteration at BASELINE configs[1] shapes (for ncu captures; not a benchmark)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

N, T = int(os.environ.get("HG_ENVS", 4096)), int(os.environ.get("HG_T", 60))
iters = int(os.environ.get("HG_ITERS", 1))
dev = "cuda:0"
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
          entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device=dev)
alg.init_storage(N, T, [705], [219], [12])
env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), dev, torch_rng_noise=False)
env.bind_storage(alg.storage, gamma=alg.gamma)
alg.torch_rng = False
obs, cobs = env.get_observations(), env.get_privileged_observations()
for it in range(iters + 1):        # iteration 0 warms up (plans, workspaces); profile the rest with --profile-from-start off
    if it == 1:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    with torch.inference_mode():
        for i in range(T):
            a = alg.act(obs, cobs)
            obs, cobs, rew, dones, infos = env.step(a)
            alg.process_env_step(rew, dones, infos)
        alg.compute_returns(cobs)
    print(alg.update())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
