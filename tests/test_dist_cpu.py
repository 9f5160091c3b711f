"""CPU test (gloo, world_size 2) of the N>1 protocol of the hot path: envs are sharded across ranks, the ONLY collective
is one SUM all-reduce per minibatch of the flat buffer [gradients | loss statistics], with every rank's gradients
pre-scaled by 1/world so that the reduced buffer holds the gradient of the global mean loss and the global KL sum
(from which every rank derives the same adaptive learning rate).  Checked here with the oracle's arithmetic: the
reduced gradient / KL / lr must equal those of one process that sees all the rows."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as po

NETS = ((64, 32), (64, 32), (32,))
N_OBS, N_PRIV, A = 47, 73, 12


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed):
    torch.manual_seed(seed)
    return po.init_params(N_OBS, N_PRIV, A, *NETS, 1.0)


def _batch(rows, seed):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    return dict(obs=rn(rows, N_OBS), critic_obs=rn(rows, N_PRIV), lin_vel_target=rn(rows, 3), actions=rn(rows, A),
                old_values=rn(rows, 1), advantages=rn(rows, 1), returns=rn(rows, 1), old_log_prob=-12 + rn(rows, 1),
                old_mu=0.3 * rn(rows, A), old_sigma=0.8 + 0.4 * torch.rand(rows, A, generator=g))


def _flat_grads(p, mb, cfg, scale):
    keys = po.state_dict_order(p)
    leaves = {k: p[k].detach().requires_grad_(True) for k in keys}
    out = po.minibatch_losses(leaves, mb, cfg)
    gl = torch.autograd.grad(out["loss"] * scale, [leaves[k] for k in keys], allow_unused=True)
    flat = torch.cat([(g if g is not None else torch.zeros_like(p[k])).reshape(-1) for k, g in zip(keys, gl)])
    rows = mb["obs"].shape[0]
    stats = torch.tensor([float(out["kl"]) * rows, float(out["value"]) * rows, float(out["surrogate"]) * rows])
    return flat, stats


def _worker(rank, world, port, rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        cfg = po.xbot_ppo_config()
        p = _make(0)                                    # identical replicas
        full = _batch(rows * world, 123)
        shard = {k: v[rank * rows:(rank + 1) * rows] for k, v in full.items()}       # this rank's envs
        flat, stats = _flat_grads(p, shard, cfg, 1.0 / world)
        buf = torch.cat([flat, stats])                  # [gradients | statistics]: ONE collective
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        kl_mean = float(buf[flat.numel()] / (rows * world))
        lr = po.adaptive_lr(cfg.learning_rate, kl_mean, cfg.desired_kl)
        q.put((rank, buf, lr))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradient_allreduce_equals_single_process():
    world, rows = 2, 96
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    cfg = po.xbot_ppo_config()
    p = _make(0)
    full = _batch(rows * world, 123)
    want_flat, want_stats = _flat_grads(p, full, cfg, 1.0)
    want = torch.cat([want_flat, want_stats])
    for rank, buf, lr in results:
        assert torch.allclose(buf[:want_flat.numel()], want_flat, rtol=1e-4, atol=1e-7), rank
        assert torch.allclose(buf[want_flat.numel():], want_stats, rtol=1e-5), rank
        assert lr == po.adaptive_lr(cfg.learning_rate, float(want_stats[0]) / (rows * world), cfg.desired_kl)
    assert torch.equal(results[0][1], results[1][1])    # replicas stay bit-identical
