"""GPU time of the three phases of one PPO iteration (rollout / returns / update) at configs[1] shapes, CUDA-event timed."""
import contextlib, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

N, T = int(os.environ.get("HG_ENVS", 4096)), int(os.environ.get("HG_T", 60))
iters = int(os.environ.get("HG_ITERS", 8))
dev = "cuda:0"
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ac = ActorCritic(705, 219, 12, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
          entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device=dev)
alg.init_storage(N, T, [705], [219], [12])
env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=N), dev, torch_rng_noise=bool(os.environ.get("HG_TORCH_NOISE")))
if not os.environ.get("HG_UNBOUND"):
    env.bind_storage(alg.storage, gamma=alg.gamma)
obs, cobs = env.get_observations(), env.get_privileged_observations()
acc = [0.0, 0.0, 0.0]
cpu = [0.0, 0.0, 0.0]
for it in range(iters + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tc = [0.0] * 4
    torch.cuda.synchronize()
    ev[0].record(); tc[0] = time.perf_counter()
    with torch.inference_mode():
        for i in range(T):
            a = alg.act(obs, cobs)
            obs, cobs, rew, dones, infos = env.step(a)
            alg.process_env_step(rew, dones, infos)
        ev[1].record(); tc[1] = time.perf_counter()
        alg.compute_returns(cobs)
    ev[2].record(); tc[2] = time.perf_counter()
    alg.update()
    ev[3].record(); tc[3] = time.perf_counter()
    upd_enq = alg.last_update_enqueue_ms
    torch.cuda.synchronize()
    if it >= 3:
        for k in range(3):
            acc[k] += ev[k].elapsed_time(ev[k + 1]) / iters
            cpu[k] += (tc[k + 1] - tc[k]) * 1e3 / iters
if os.environ.get("HG_NOSYNC"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for it in range(iters):
        with torch.inference_mode():
            for i in range(T):
                a = alg.act(obs, cobs)
                obs, cobs, rew, dones, infos = env.step(a)
                alg.process_env_step(rew, dones, infos)
            alg.compute_returns(cobs)
        alg.update()
    e1.record()
    torch.cuda.synchronize()
    print(f"back-to-back: {e0.elapsed_time(e1) / iters:.3f} ms / iteration")
print(f"envs {N} T {T} PDL={os.environ.get('HG_PDL', 'default')}: rollout {acc[0]:.3f} ms  returns {acc[1]:.3f} ms  update {acc[2]:.3f} ms  total {sum(acc):.3f} ms"
      f" | host enqueue: {cpu[0]:.2f} / {cpu[1]:.2f} / {cpu[2]:.2f} ms (update enqueue alone {upd_enq:.2f} ms)")
