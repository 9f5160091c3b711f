"""Worker of tests/test_dist_gpu.py, one process per rank (launched by torch.distributed.run).

Checks the PRODUCT's N > 1 path -- `PPO.update()` with `world_size` ranks, envs sharded, ONE all-reduce of
[gradients | loss statistics] per minibatch -- against a single-process update over the union of the shards:
  1. replicas built under DIFFERENT torch seeds are made identical by `PPO.sync_replicas()` (broadcast from rank 0),
  2. the sharded update equals the single-process update (per-minibatch losses, KL, identical lr decisions, parameters),
  3. replicas stay bit-identical after the update,
  4. BASELINE configs[3] in small: the chained env + learner loop (fused env step -> act -> store -> GAE -> update) for two
     iterations per rank keeps the replicas bit-identical while every rank explores with its own Philox stream.
Backend: NCCL with one GPU per rank when the box has enough GPUs, otherwise gloo with every rank on cuda:0 (the all-reduce
call in ppo.py is the same; only the transport differs)."""
import contextlib
import datetime
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic
from humanoid_gym_with_comments_b200.envs import SyntheticXBotEnv, XBotMathConfig

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
nccl = torch.cuda.device_count() >= world
dev = f"cuda:{local}" if nccl else "cuda:0"
torch.cuda.set_device(dev)
_timeout = datetime.timedelta(seconds=int(os.environ.get("HG_DIST_TIMEOUT_S", "120")))     # a lost peer raises instead of hanging
if nccl:
    dist.init_process_group("nccl", device_id=torch.device(dev), timeout=_timeout)
else:
    dist.init_process_group("gloo", timeout=_timeout)
if rank == 0:
    print("process group up:", "nccl" if nccl else "gloo", flush=True)
NL, T, A = 512, 8, 12            # envs per rank
N = NL * world


def make(n_envs, seed, ws=None):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, A, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
    # ws = 1: a single-process replica built by ONE rank -- it must not join the group's collectives (the constructor's
    # broadcast would pair up with whatever the other rank does next and deadlock)
    PPO.dist_override = (ws, 0) if ws is not None else None
    try:
        alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
                  entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device=dev)
    finally:
        PPO.dist_override = None
    alg.init_storage(n_envs, T, [705], [219], [A])
    return alg


def gather(t):
    t = t.contiguous() if nccl else t.detach().cpu().contiguous()      # gloo has no CUDA all_gather
    got = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(got, t)
    return got


def all_equal(t):
    got = gather(t)
    return all(torch.equal(got[0], x) for x in got)


g = torch.Generator().manual_seed(42)                       # the same full data set on every rank
full = dict(obs=torch.randn(T, N, 705, generator=g), priv=torch.randn(T, N, 219, generator=g),
            actions=torch.randn(T, N, A, generator=g), mu=0.3 * torch.randn(T, N, A, generator=g),
            sigma=torch.ones(T, N, A), logp=-17 + torch.randn(T, N, 1, generator=g), values=torch.randn(T, N, 1, generator=g),
            returns=torch.randn(T, N, 1, generator=g), adv=torch.randn(T, N, 1, generator=g))


def fill(alg, sl):
    st = alg.storage
    for t in range(T):
        st.ingest(full["obs"][t, sl].to(dev), full["priv"][t, sl].to(dev), t)
    for name, key in (("actions", "actions"), ("mu", "mu"), ("sigma", "sigma"), ("actions_log_prob", "logp"),
                      ("values", "values"), ("returns", "returns"), ("advantages", "adv")):
        getattr(st, name).copy_(full[key][:, sl].to(dev))


# 1. different seeds per rank: the constructor's broadcast must make the replicas identical (weights of rank 0 = seed 0)
alg = make(NL, seed=rank)
assert alg.world_size == world and alg.rank == rank
assert all_equal(alg.actor_critic.flat.p), "sync_replicas() did not equalise the replicas"
seeds = gather(torch.tensor([alg._seed & (2 ** 62 - 1)], dtype=torch.int64, device=dev))
assert len({int(s) for s in seeds}) == world, "every rank needs its own Philox key"

# 2. sharded update: rank r owns envs [r*NL, (r+1)*NL)
fill(alg, slice(rank * NL, (rank + 1) * NL))
perms = [torch.randperm(NL * T, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
alg.perm_override = perms[rank].to(dev)
out = alg.update()
log = alg.last_update_log.clone()
flat = alg.actor_critic.flat.p.clone()
ok = all_equal(flat)                                        # 3. replicas stay bit-identical

if rank == 0:
    # single-process equivalent over all N envs: minibatch i = union over ranks of their minibatch i
    ref = make(N, seed=0, ws=1)
    fill(ref, slice(0, N))
    mb = NL * T // 4
    rows = []
    for i in range(4):
        for r in range(world):
            loc = perms[r][i * mb:(i + 1) * mb]
            t_idx, n_loc = loc // NL, loc % NL
            rows.append(t_idx * N + r * NL + n_loc)
    ref.perm_override = torch.cat(rows).to(dev)
    rout = ref.update()
    rlog = ref.last_update_log
    dp = float((ref.actor_critic.flat.p - flat).abs().max())
    print("backend:", "nccl" if nccl else "gloo (all ranks on cuda:0)", "| replicas identical:", ok)
    print("sharded  :", [f"{x:.6f}" for x in out[:2]], out[3], "lr", alg.learning_rate)
    print("single   :", [f"{x:.6f}" for x in rout[:2]], rout[3], "lr", ref.learning_rate)
    print("max |loss log diff| (value, surrogate, lin-vel, kl, entropy):", (log[:, :5] - rlog[:, :5]).abs().max(0).values.tolist())
    print("grad-norm rel diff:", float(((log[:, 6] - rlog[:, 6]).abs() / rlog[:, 6]).max()), "max |param diff|:", dp)
    assert ok
    assert torch.allclose(log[:, :5], rlog[:, :5], rtol=1e-4, atol=1e-6)
    assert torch.equal(log[:, 5], rlog[:, 5])                  # identical adaptive-lr decisions
    assert dp <= 2.5 * ref.learning_rate * 8
    del ref

# 4. configs[3] in small: chained env + learner loop, two iterations
torch.manual_seed(1000 + rank)
env = SyntheticXBotEnv(XBotMathConfig.for_dof(12, num_envs=NL), dev, pool=4, seed=rank, torch_rng_noise=False)
alg.torch_rng = False
alg.perm_override = None
env.bind_storage(alg.storage, gamma=alg.gamma)
obs, cobs = env.get_observations(), env.get_privileged_observations()
first_actions = None
for it in range(2):
    with torch.inference_mode():
        for i in range(T):
            a = alg.act(obs, cobs)
            if first_actions is None:
                first_actions = a.clone()
            obs, cobs, rew, dones, infos = env.step(a)
            alg.process_env_step(rew, dones, infos)
        alg.compute_returns(cobs)
    losses = alg.update()
    assert all(torch.isfinite(torch.tensor(float(x))) for x in losses)
ok2 = all_equal(alg.actor_critic.flat.p) and all_equal(alg.actor_critic.flat.m) and all_equal(alg.optimizer.opt_state)
acts = gather(first_actions)
independent = not torch.equal(acts[0], acts[1])
if rank == 0:
    print("chained env + learner loop: replicas identical:", ok2, "| ranks explore independently:", independent,
          "| lr", alg.learning_rate)
    assert ok2 and independent
    print("DIST CHECK OK")
dist.barrier()
dist.destroy_process_group()
