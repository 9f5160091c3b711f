import time, torch, contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig
dev="cuda:0"; N,T=int(os.environ.get("HG_ENVS",4096)),60
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ac = ActorCritic(705,219,12,actor_hidden_dims=[512,256,128],critic_hidden_dims=[768,256,128])
alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, gamma=0.994, lam=0.9, entropy_coef=0.001, learning_rate=1e-5, schedule="adaptive", device=dev)
alg.init_storage(N,T,[705],[219],[12]); alg.torch_rng=False
env = SyntheticXBotEnv(XBotMathConfig.for_dof(12,num_envs=N), dev, torch_rng_noise=False); env.bind_storage(alg.storage)
obs, cobs = env.get_observations(), env.get_privileged_observations()
def rollout():
    global obs, cobs
    with torch.inference_mode():
        for i in range(T):
            a = alg.act(obs, cobs); obs, cobs, r, d, info = env.step(a); alg.process_env_step(r, d, info)
        alg.compute_returns(cobs)
for it in range(3):
    rollout(); alg.update()
torch.cuda.synchronize()
for it in range(3):
    t0=time.perf_counter(); rollout(); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    alg.update(); t3=time.perf_counter(); torch.cuda.synchronize(); t4=time.perf_counter()
    print(f"N={N} rollout: cpu enqueue {1e3*(t1-t0):.2f} ms, gpu done {1e3*(t2-t0):.2f} ms; update: cpu {1e3*(t3-t2):.2f} ms, done {1e3*(t4-t2):.2f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); rollout(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
