"""Run under torchrun on W GPUs: the W-rank update (envs sharded, one NCCL all-reduce of [grads | stats] per minibatch)
must equal a single-process update over the union of the shards with the matching permutation.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from humanoid_gym_with_comments_b200.algo.ppo import PPO, ActorCritic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
NL, T, A = 512, 8, 12            # envs per rank
N = NL * world


def make(n_envs, ws):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        ac = ActorCritic(705, 219, A, actor_hidden_dims=[512, 256, 128], critic_hidden_dims=[768, 256, 128])
    alg = PPO(ac, num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9, value_loss_coef=1.0,
              entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0, schedule="adaptive", desired_kl=0.01, device=dev)
    alg.world_size = ws
    alg.init_storage(n_envs, T, [705], [219], [A])
    return alg


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


# sharded run: rank r owns envs [r*NL, (r+1)*NL)
alg = make(NL, world)
fill(alg, slice(rank * NL, (rank + 1) * NL))
perms = [torch.randperm(NL * T, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
alg.perm_override = perms[rank].to(dev)
out = alg.update()
log = alg.last_update_log.clone()
flat = alg.actor_critic.flat.p.clone()
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
ok = all(torch.equal(gathered[0], x) for x in gathered)      # replicas stay bit-identical

if rank == 0:
    # single-process equivalent over all N envs: minibatch i = union over ranks of their minibatch i
    ref = make(N, 1)
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
    print("replicas identical:", ok)
    print("sharded  :", [f"{x:.6f}" for x in out[:2]], out[3], "lr", alg.learning_rate)
    print("single   :", [f"{x:.6f}" for x in rout[:2]], rout[3], "lr", ref.learning_rate)
    print("max |loss log diff| (value, surrogate, lin-vel, kl, entropy):", (log[:, :5] - rlog[:, :5]).abs().max(0).values.tolist())
    print("grad-norm rel diff:", float(((log[:, 6] - rlog[:, 6]).abs() / rlog[:, 6]).max()), "max |param diff|:", dp)
    assert ok
    assert torch.allclose(log[:, :5], rlog[:, :5], rtol=1e-4, atol=1e-6)
    assert torch.equal(log[:, 5], rlog[:, 5])
    assert dp <= 2.5 * ref.learning_rate * 8
    print("DIST CHECK OK")
dist.barrier()
dist.destroy_process_group()
