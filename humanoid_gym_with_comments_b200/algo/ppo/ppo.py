"""`PPO` with the reference's constructor, attributes and methods (reference: humanoid/algo/ppo/ppo.py:42-226),
every tensor operation on sm_100a kernels and no host synchronisation inside rollout or update (one device->host
read of the loss log at the end of `update()`).

Per rollout step (`act` + `process_env_step`, ppo.py:116-138), generic path: 2 ingest launches, one grouped tcgen05 GEMM
launch per hidden depth (actor + critic together), 1 fused head launch (output layers, Gaussian sample, log-prob, value,
written straight into the storage slot), 1 reward/done launch.  With a bound env (`SyntheticXBotEnv.bind_storage`) and
three-hidden-layer nets: the layer-1 launch + ONE launch for layer 2, layer 3, output layers and sampling; ingest and
reward/done bookkeeping are done by the env kernel.
Per minibatch (`update`, ppo.py:150-218): grouped forward GEMMs, 1 fused head+loss+head-backward launch, grouped
dW / dX GEMMs, optional NCCL all-reduce of the flat gradient arena (+KL slot), 2 optimiser launches.
"""
from __future__ import annotations

import ctypes as C

import os
import time

import torch
import torch.optim as optim

from ... import _lib
from ..._lib import HgAdamArgs, HgHeadNet, HgLossArgs, check, current_stream, nvtx_range, ptr
from .actor_critic import ActorCritic
from .engine import LD_ALIGN
from .engine import round_up as _round_up
from .rollout_storage import RolloutStorage


class FusedAdam(optim.Adam):
    """torch.optim.Adam-shaped facade (state_dict / load_state_dict / param_groups) over the fused
    clip+Adam kernel: `exp_avg` / `exp_avg_sq` of every parameter are views of the flat moment arenas."""

    def __init__(self, actor_critic: ActorCritic, lr: float):
        super().__init__(actor_critic.parameters(), lr=lr)
        self._ac = actor_critic
        self._flat = None
        self._load_version = 0
        # device scalars: [0] lr [1] step [2] grad sumsq [3] clip coef [4] step size [5] sqrt(1-b2^t) [6] grad norm
        self.opt_state = None

    def bind(self):
        flat = self._ac.flat
        if flat is self._flat:
            return
        old = self._flat
        self._flat = flat
        lr = float(self.param_groups[0]["lr"])
        step = 0.0
        if old is not None and old.count == flat.count:
            flat.m.copy_(old.m)
            flat.v.copy_(old.v)
            step = float(self.opt_state[1].item())
        self.opt_state = torch.zeros(8, dtype=torch.float64, device=flat.device)
        self.opt_state[0] = lr
        self.opt_state[1] = step
        for p in flat.params:
            m, v = flat.moment_views(p)
            self.state[p] = {"step": torch.tensor(step), "exp_avg": m, "exp_avg_sq": v}

    def sync_host(self):
        """Mirror device-side lr / step into param_groups and state (one small D2H copy)."""
        if self.opt_state is None:
            return
        lr, step = self.opt_state[:2].tolist()
        for g in self.param_groups:
            g["lr"] = lr
        for st in self.state.values():
            st["step"] = torch.tensor(float(step))

    def set_lr(self, lr: float):
        for g in self.param_groups:
            g["lr"] = lr
        if self.opt_state is not None:
            self.opt_state[0] = lr

    def state_dict(self):
        self.sync_host()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._load_version += 1
        flat = self._ac.flat
        self._flat = None
        loaded = {id(p): self.state.get(p) for p in flat.params}
        step = 0.0
        with torch.no_grad():
            for p in flat.params:
                st = loaded[id(p)]
                m, v = flat.moment_views(p)
                if st:
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                    step = float(st["step"])
        self._flat = flat
        self.opt_state = torch.zeros(8, dtype=torch.float64, device=flat.device)
        self.opt_state[0] = float(self.param_groups[0]["lr"])
        self.opt_state[1] = step
        for p in flat.params:
            m, v = flat.moment_views(p)
            self.state[p] = {"step": torch.tensor(step), "exp_avg": m, "exp_avg_sq": v}

    def zero_grad(self, set_to_none: bool = False):
        self._ac.flat.g.zero_()

    def fused_step(self, *, rows_global, max_grad_norm, adaptive, desired_kl, loss_log, log_slot):
        self.bind()
        flat = self._flat
        g = self.param_groups[0]
        args = HgAdamArgs(
            params=ptr(flat.p), grads=ptr(flat.g), exp_avg=ptr(flat.m), exp_avg_sq=ptr(flat.v), count=flat.count,
            opt_state=ptr(self.opt_state), loss_log=ptr(loss_log), log_slot=log_slot, rows_global=rows_global,
            beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"],
            max_grad_norm=float(max_grad_norm) if max_grad_norm is not None else 3.0e38,
            adaptive=1 if adaptive else 0, desired_kl=float(desired_kl) if desired_kl is not None else 0.0,
            lr_min=1e-5, lr_max=1e-2, shadows_host=flat.shadow_array, num_shadows=len(flat.shadow_array),
            zero_grads=1)
        check(_lib.load().hg_clip_adam_step(C.byref(args), current_stream()))

    def step(self, closure=None):
        """Plain Adam step on whatever is in the gradient arena (no clipping, fixed lr)."""
        self.bind()
        self.opt_state[0] = float(self.param_groups[0]["lr"])
        self.fused_step(rows_global=1, max_grad_norm=None, adaptive=False, desired_kl=None, loss_log=None, log_slot=0)


class PPO:
    actor_critic: ActorCritic
    # (world_size, rank) to use instead of torch.distributed's while this is set: lets one rank of a running process group
    # build a single-process replica (a PPO that must not join the group's collectives), e.g. the 2-rank parity test
    dist_override = None

    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0,
                 use_clipped_value_loss=True, schedule="fixed", desired_kl=0.01, device='cpu', sym_loss=False,
                 obs_permutation=None, act_permutation=None, frame_stack=0, sym_coef=1.0, base_lin_vel_coef=1.0):
        if torch.device(device).type != "cuda":
            raise _lib.HgError("PPO(device=...) must be a CUDA device: this implementation has no CPU fallback")
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate

        # PPO components
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        self.actor_critic.engine()
        self.storage = None  # initialized later
        self.optimizer = FusedAdam(self.actor_critic, lr=learning_rate)
        self.optimizer.bind()
        self.transition = RolloutStorage.Transition()

        # PPO parameters
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.base_lin_vel_coef = base_lin_vel_coef
        self.sym_loss = sym_loss
        self.sym_coef = sym_coef
        self.last_sym_loss = 0.0
        if self.sym_loss:      # signed permutations of ppo.py:95-105 kept as (source column, sign) instead of dense matrices
            import numpy as np
            if obs_permutation is None or act_permutation is None or frame_stack <= 0:
                raise ValueError("sym_loss=True needs obs_permutation, act_permutation and frame_stack (ppo.py:59-63)")
            stack = [np.sign(q) * (abs(q) + i * len(obs_permutation)) for i in range(frame_stack) for q in obs_permutation]
            # the reference multiplies by dense [O,O] / [A,A] matrices and fails on a shape mismatch; the kernels index,
            # so the same mistakes are rejected here instead of becoming out-of-range reads
            O, A_ = self.actor_critic.num_actor_obs, self.actor_critic.num_actions
            if len(stack) != O:
                raise ValueError(f"len(obs_permutation) * frame_stack = {len(stack)} must equal num_actor_obs = {O}")
            if len(act_permutation) != A_:
                raise ValueError("act_permutation must have one entry per action")
            if max(int(abs(q)) for q in stack) >= O or max(int(abs(q)) for q in act_permutation) >= A_:
                raise ValueError("permutation index out of range")
            if len({int(abs(q)) for q in act_permutation}) != A_ or len({int(abs(q)) for q in stack}) != O:
                raise ValueError("obs_permutation / act_permutation must be true (signed) permutations: the symmetry "
                                 "kernels scatter d/d(mirror) through the index, duplicates would drop gradient")
            dev = self.device
            as_src = lambda perm: torch.tensor([int(abs(q)) for q in perm], dtype=torch.int32, device=dev)
            as_sign = lambda perm: torch.tensor([float(np.sign(q)) for q in perm], dtype=torch.float32, device=dev)
            self._obs_src, self._obs_sign = as_src(stack), as_sign(stack)
            self._act_src, self._act_sign = as_src(act_permutation), as_sign(act_permutation)
            self._sym_ws = None

        # implementation switches
        self.torch_rng = True        # draw eps / burn the update-time sample() through torch's generator (seed parity)
        dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
        self.world_size = torch.distributed.get_world_size() if dist_on else 1
        self.rank = torch.distributed.get_rank() if dist_on else 0
        if PPO.dist_override is not None:
            self.world_size, self.rank = PPO.dist_override
        self.last_update_log = None  # [epochs*minibatches, 8] per-minibatch statistics of the last update()
        self._perm_bufs = None
        self._lv_targets = None
        self._lr_synced = learning_rate
        # test hooks: replay externally drawn randomness (parity tests feed the oracle's eps / permutation)
        self._lib = _lib.load()
        self._slot_plans, self._plans_owner, self._plans_storage, self._plans_ring = {}, None, None, None
        # Philox key of the action-noise stream: torch's seed mixed with the rank, so that data-parallel replicas (built under
        # the SAME torch seed to get identical weights) explore independently -- (seed, step, local row) alone would give
        # every rank bit-identical eps for the same local env row
        self._seed = self.actor_critic.philox_seed(self.rank)
        self._replica_versions = None
        self.allreduce_kind = "none"
        if self.world_size > 1:
            self.sync_replicas()
            self._setup_allreduce()
        self._gamma_f = float(gamma)
        self.eps_override = None     # callable(step) -> [N, A] standard-normal draws
        self.fuse_rollout_head = os.environ.get("HG_FUSE_HEAD", "1") != "0"   # last hidden layer + output layer + sampling in one launch
        self.perm_override = None    # int64 [num_mini_batches * mini_batch_size] permutation

    def sync_replicas(self):
        """world_size > 1: the update all-reduces gradients only, so parameters, Adam moments and lr / step must be
        identical on every rank.  Broadcast them from rank 0 (construction, and again after either load_state_dict) and
        re-cast the bf16 weight shadows."""
        flat, opt = self.actor_critic.flat, self.optimizer
        opt.bind()
        for t in (flat.p, flat.m, flat.v, opt.opt_state):
            torch.distributed.broadcast(t, src=0)
        flat.refresh_shadows()
        opt.sync_host()
        self.learning_rate = self._lr_synced = float(opt.param_groups[0]["lr"])
        self._replica_versions = (self.actor_critic._load_version, opt._load_version)

    def _setup_allreduce(self):
        """The path's one collective: SUM of [gradients | loss statistics] over the ranks, once per minibatch.
        Optional transport (HG_ALLREDUCE=symm): the gradient arena lives in symmetric memory (every rank maps its peers'
        arenas over NVLink) and the reduction runs as torch's two-shot multimem all-reduce -- `multimem.ld_reduce` sums a slice
        inside the NVSwitch, `multimem.st` broadcasts it; without multicast support (or HG_ALLREDUCE=two_shot) the peer-to-peer
        two-shot kernel.  Any failure to set symmetric memory up falls back to `dist.all_reduce`."""
        self.allreduce_kind = "nccl"
        # Default: NCCL.  Measured on 2 x B200 (tools/bench_allreduce.py, 4.13 MB): NCCL 37.9 us, two-shot multimem 46.5 us,
        # peer-to-peer two-shot 28.0 us -- no gain worth a second code path at N = 2, and the 8-GPU run of the same
        # microbenchmark did not finish within its time limit (not investigated), so the symmetric-memory transport stays
        # opt-in (HG_ALLREDUCE=symm | two_shot) until it has been qualified at N = 8.
        want = os.environ.get("HG_ALLREDUCE", "nccl")
        if want == "nccl" or torch.distributed.get_backend() != "nccl":
            return
        try:
            import torch.distributed._symmetric_memory as symm
            flat = self.actor_critic.flat
            group = torch.distributed.group.WORLD
            n = flat.g.numel()
            padded = _round_up(n, 4 * 32 * self.world_size)      # every rank's slice a whole number of 16-byte chunks
            buf = symm.empty(padded, dtype=torch.float32, device=flat.g.device)
            hdl = symm.rendezvous(buf, group)
            buf.zero_()
            flat.rebind_grads(buf[:n])
            self._ar_buf, self._ar_group = buf, group.group_name
            mc = getattr(hdl, "has_multicast_support", False)
            multicast = bool(mc() if callable(mc) else mc) and want != "two_shot"
            self.allreduce_kind = "symm_multimem" if multicast else "symm_two_shot"
            self._ar_op = torch.ops.symm_mem.multimem_all_reduce_ if multicast else torch.ops.symm_mem.two_shot_all_reduce_
            self._ar_op(self._ar_buf, "sum", self._ar_group)     # first call: any lazy set-up happens here, on zeros
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001 -- any set-up problem: the NCCL all-reduce is always available
            self.allreduce_kind = "nccl"
            self._ar_fallback_reason = repr(e)
        # the choice must be the same on every rank (a collective met by a different collective would hang)
        kinds = {"nccl": 0, "symm_two_shot": 1, "symm_multimem": 2}
        agreed = torch.tensor([kinds[self.allreduce_kind]], device=self.actor_critic.flat.g.device, dtype=torch.int32)
        lo = agreed.clone()
        torch.distributed.all_reduce(agreed, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        if int(agreed.item()) != int(lo.item()):
            self.allreduce_kind = "nccl"

    def _allreduce_grads(self):
        g = self.actor_critic.flat.g
        if self.allreduce_kind == "nccl" or g.data_ptr() != self._ar_buf.data_ptr():     # (a rebuilt arena is no longer symmetric)
            torch.distributed.all_reduce(g)
        else:
            self._ar_op(self._ar_buf, "sum", self._ar_group)

    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.actor_critic.eval()   # the reference calls a non-existent .test() (ppo.py:111)

    def train_mode(self):
        self.actor_critic.train()

    # ------------------------------------------------------------------ rollout
    def _slot_plan(self, t):
        """Everything `act` / `process_env_step` need for storage slot t, prepared once: the grouped-GEMM plans
        (TMA tensor maps encoded, tile schedule laid out), the head launch arguments and the slot views."""
        ac, st = self.actor_critic, self.storage
        eng = ac.engine()
        ring = st.ring_enabled
        if self._plans_owner is not eng or self._plans_storage is not st or self._plans_ring != (st._ring_obs if ring else None):
            self._slot_plans = {}
            self._plans_owner, self._plans_storage, self._plans_ring = eng, st, (st._ring_obs if ring else None)
        plan = self._slot_plans.get(t)
        if plan is None:
            N, A = st.num_envs, ac.num_actions
            if ring:      # layer 1 reads the history ring's bf16 window of row t in place (TMA over the strided view)
                o16, p16 = st._ring_obs.view16(t), st._ring_priv.view16(t)
                in_cols = {"actor": st._ring_obs.K16, "critic": st._ring_priv.K16}
                frames = {"actor": (st._ring_obs.width, st._ring_obs.pitch16), "critic": (st._ring_priv.width, st._ring_priv.pitch16)}
            else:
                _, _, o16, p16 = st.slot(t)
                in_cols, frames = {"actor": st.obs_cols, "critic": st.priv_cols}, None
            ws = ac.workspace(N)
            la, lc = ac.net_actor.linears[-1], ac.net_critic.linears[-1]
            heads = None
            if self.fuse_rollout_head and eng.can_fuse_heads((ac.net_actor, ac.net_critic), (A, 1)):
                heads = {"actor": _lib.HgFusedHead(kind=_lib.HEAD_POLICY, n_out=A, W=ptr(la.weight.data), b=ptr(la.bias.data),
                                                   out=ptr(st.mu[t]), std=ptr(ac.std.data), actions=ptr(st.actions[t]),
                                                   sigma=ptr(st.sigma[t]), log_prob=ptr(st.actions_log_prob[t])),
                         "critic": _lib.HgFusedHead(kind=_lib.HEAD_LINEAR, n_out=1, W=ptr(lc.weight.data), b=ptr(lc.bias.data),
                                                    out=ptr(st.values[t]))}
            gemms = eng.forward_hidden((ac.net_actor, ac.net_critic), {"actor": o16, "critic": p16}, in_cols, ws, N,
                                       make_plans=True, heads=heads, frames=frames)
            if heads is not None:
                gemms = (gemms[:-1], gemms[-1])          # (plain plans, fused last-layer + head plan)
            ha, hc = ws.h["actor"][-1], ws.h["critic"][-1]
            views = dict(actions=st.actions[t], values=st.values[t], log_prob=st.actions_log_prob[t].view(-1),
                         mu=st.mu[t], sigma=st.sigma[t], rewards=st.rewards[t].view(-1))
            head_args = (ptr(ha), ha.stride(0), ha.shape[1], ptr(la.weight.data), ptr(la.bias.data), ptr(ac.std.data),
                         ptr(hc), hc.stride(0), hc.shape[1], ptr(lc.weight.data), ptr(lc.bias.data))
            head_out = (N, A, ptr(st.actions[t]), ptr(st.mu[t]), ptr(st.sigma[t]), ptr(st.actions_log_prob[t]),
                        ptr(st.values[t]))
            store_ptrs = (ptr(st.values[t]), ptr(st.rewards[t]), ptr(st.dones[t]))
            plan = (gemms, head_args, head_out, views, store_ptrs)
            self._slot_plans[t] = plan
        return plan

    def act(self, obs, critic_obs):
        ac, st = self.actor_critic, self.storage
        t = st.step
        if t >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        if critic_obs is None:
            critic_obs = obs
        gemms, head_args, head_out, views, _ = self._slot_plan(t)
        if not st.holds(obs, critic_obs, t):          # a bound env already wrote this slot and its bf16 shadow
            st.ingest(obs, critic_obs, t)
        if self.eps_override is not None:
            eps = self.eps_override(t).to(st.actions.device).float().contiguous()
        else:
            eps = torch.randn(head_out[0], head_out[1], device=st.actions.device) if self.torch_rng else None
        ac._rng_counter += 1
        if isinstance(gemms, tuple):                  # last hidden layer, output layers and sampling in one launch
            chain = views.get("chain")
            if chain is None:                         # (plan pointers as a C array: the whole forward is one binding call)
                chain = views["chain"] = ((C.c_void_p * len(gemms[0]))(*gemms[0]), len(gemms[0]), gemms[1])
            if self._lib.hg_rollout_forward(chain[0], chain[1], chain[2], ptr(eps), self._seed, ac._rng_counter, current_stream()) != 0:
                check(-1)
        else:
            ac._engine.run_plans(gemms)
            check(self._lib.hg_policy_head_rollout(*head_args, ptr(eps), self._seed, ac._rng_counter, *head_out,
                                                   current_stream()))
        tr = self.transition
        tr.actions = views["actions"]
        tr.values = views["values"]
        tr.actions_log_prob = views["log_prob"]
        tr.action_mean = views["mu"]
        tr.action_sigma = views["sigma"]
        tr.observations = obs
        tr.critic_observations = critic_obs
        return tr.actions

    def process_env_step(self, rewards, dones, infos):
        st = self.storage
        t = st.step
        if t >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        stored = getattr(st, "env_stored", None)
        if stored is not None and stored == (t, rewards.data_ptr(), dones.data_ptr(), self._gamma_f):
            # a bound env (SyntheticXBotEnv.bind_storage(storage, gamma)) already wrote this transition's record
            st.env_stored = None
            st.step = t + 1
            self.transition.clear()
            self.actor_critic.reset(dones)
            return
        val_ptr, rew_ptr, done_ptr = self._slot_plan(t)[4]
        if not (rewards.is_cuda and dones.is_cuda):
            _lib.require_cuda(rewards, dones)
        if rewards.dtype != torch.float32 or rewards.dim() != 1 or not rewards.is_contiguous():
            rewards = rewards.reshape(-1).float().contiguous()
        d8 = dones
        if d8.dtype not in (torch.bool, torch.uint8) or d8.dim() != 1 or not d8.is_contiguous():
            d8 = d8.reshape(-1)
            d8 = (d8 if d8.dtype in (torch.bool, torch.uint8) else (d8 != 0)).contiguous()
        to8 = infos.get('time_outs') if infos else None
        if to8 is not None and (to8.dtype not in (torch.bool, torch.uint8) or to8.dim() != 1 or not to8.is_contiguous()
                                or not to8.is_cuda):
            to8 = to8.reshape(-1).to(self.device)
            to8 = (to8 if to8.dtype in (torch.bool, torch.uint8) else (to8 != 0)).contiguous()
        check(self._lib.hg_store_rewards_dones(ptr(rewards), val_ptr, ptr(to8), ptr(d8), self._gamma_f, st.num_envs,
                                               rew_ptr, done_ptr, current_stream()))
        st.step = t + 1
        self.transition.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs):
        with nvtx_range("hg.ppo.compute_returns"):
            return self._compute_returns(last_critic_obs)

    def _compute_returns(self, last_critic_obs):
        ac, st = self.actor_critic, self.storage
        eng = ac.engine()
        N = last_critic_obs.shape[0]
        T = st.num_transitions_per_env
        k_in, frames = last_critic_obs.shape[1], None
        if st.ring_enabled and N == st.num_envs and last_critic_obs.data_ptr() == st._ring_priv.ptr32(st.step):
            rp = st._ring_priv                        # the critic window of the last row, read in place from the ring
            c16, k_in, frames = rp.view16(st.step), rp.K16, {"critic": (rp.width, rp.pitch16)}
        elif st._priv_all is not None and N == st.num_envs and last_critic_obs.data_ptr() == st._priv_all[T].data_ptr():
            c16 = st.slot(T)[3]                       # written (with its shadow) by a bound env
        else:
            c16 = ac._to_bf16(last_critic_obs, "last_critic")
        ws = ac.workspace(N)
        eng.forward_hidden((ac.net_critic,), {"critic": c16}, {"critic": k_in}, ws, N, frames=frames)
        last_values = torch.empty(N, 1, device=st.values.device)
        ac._head_rollout(ws, N, False, None, dict(values=last_values))
        st.compute_returns(last_values, self.gamma, self.lam)

    # ------------------------------------------------------------------ update
    def _head_net(self, net, ws, flat) -> HgHeadNet:
        last = net.linears[-1]
        top = net.linears[-2]
        h = ws.h[net.name][-1]
        dz = ws.dz[net.name][-1]
        return HgHeadNet(h=ptr(h), ld_h=h.stride(0), H=h.shape[1], n_out=last.out_features,
                         W=ptr(last.weight.data), b=ptr(last.bias.data), gW=ptr(flat.grad_view(last.weight)),
                         gb=ptr(flat.grad_view(last.bias)), dZ=ptr(dz), ld_dz=dz.stride(0),
                         gb_hidden=ptr(flat.grad_view(top.bias)))

    def update(self):
        with nvtx_range("hg.ppo.update"):
            return self._update()

    def _update(self):
        ac, st = self.actor_critic, self.storage
        eng = ac.engine()
        flat = ac.flat
        lib = _lib.load()
        t_host0 = time.perf_counter()
        self.optimizer.bind()
        if self.world_size > 1 and self._replica_versions != (ac._load_version, self.optimizer._load_version):
            self.sync_replicas()                                # a checkpoint was loaded since the last broadcast
        if abs(self.learning_rate - self._lr_synced) > 0:       # user changed alg.learning_rate by hand
            self.optimizer.set_lr(self.learning_rate)
        nmb, nep = self.num_mini_batches, self.num_learning_epochs
        A = ac.num_actions
        indices, M = st.draw_permutation(nmb)                    # once per update, reused by every epoch
        if self.perm_override is not None:
            indices = self.perm_override.to(indices.device).long().contiguous()
        total = nmb * M
        ring = st.ring_enabled
        # gather the bf16 observation rows of all minibatches once (both epochs reuse them)
        if ring:
            ro, rp = st._ring_obs, st._ring_priv
            k_obs, k_priv = ro.K16, rp.K16
            ld_o, ld_p = _round_up(k_obs, LD_ALIGN), _round_up(k_priv, LD_ALIGN)
            frames = {"actor": (ro.width, ro.pitch16), "critic": (rp.width, rp.pitch16), "base_lin_vel": (ro.width, ro.pitch16)}
        else:
            k_obs, k_priv, ld_o, ld_p, frames = st.obs_cols, st.priv_cols, st.obs_ld, st.priv_ld, None
        if self._perm_bufs is None or self._perm_bufs[0].shape != (total, ld_o) or (self._perm_bufs[1] is not None and self._perm_bufs[1].shape[1] != ld_p):
            self._perm_bufs = (torch.zeros(total, ld_o, dtype=torch.bfloat16, device=indices.device),
                               torch.zeros(total, ld_p, dtype=torch.bfloat16, device=indices.device) if st._has_priv else None)
        xperm, pperm = self._perm_bufs
        if ring:
            # rows straight out of the archive ring: window of (t, n), frames older than the row's last reset zero, bf16 at the
            # frame pitch; the lin-vel regression targets critic_obs[:, 53:56] (rollout_storage.py:164) in storage order
            T_, N_ = st.num_transitions_per_env, st.num_envs
            if self._lv_targets is None or self._lv_targets.shape[0] != T_ * N_:
                self._lv_targets = torch.empty(T_ * N_, 4, device=indices.device)
            fv = ptr(st.ring_first_valid)
            check(lib.hg_ring_gather(C.byref(ro.c), N_, fv, ptr(indices), total, ptr(xperm), ld_o, 1, None, 0, current_stream()))
            check(lib.hg_ring_gather(C.byref(rp.c), N_, fv, ptr(indices), total, ptr(pperm), ld_p, 1, None, 0, current_stream()))
            check(lib.hg_ring_gather(C.byref(rp.c), N_, fv, None, T_ * N_, None, 0, 0, ptr(self._lv_targets), 53, current_stream()))
            cobs_ptr, ld_cobs, lv_col = ptr(self._lv_targets), 4, 0
        else:
            check(lib.hg_gather_rows_bf16(ptr(st.obs16), st.obs_ld, ptr(indices), total, ptr(xperm), current_stream()))
            if pperm is not None:
                check(lib.hg_gather_rows_bf16(ptr(st.priv16), st.priv_ld, ptr(indices), total, ptr(pperm), current_stream()))
            cobs = st.privileged_observations if st._has_priv else st.observations
            cobs_ptr, ld_cobs, lv_col = ptr(cobs), cobs.shape[-1], 53
        if pperm is None:
            pperm = xperm
        ws = ac.workspace(M, with_grads=True)
        nets = (ac.net_actor, ac.net_critic, ac.net_linvel)
        n_updates = nep * nmb
        self._sym_slot = None
        loss_log = torch.zeros(n_updates, 8, dtype=torch.float64, device=indices.device)
        adaptive = self.desired_kl is not None and self.schedule == 'adaptive'
        slot = 0
        for epoch in range(nep):
            for i in range(nmb):
                if _lib.NVTX:
                    torch.cuda.nvtx.range_push(f"hg.minibatch e{epoch} m{i}")
                x = xperm[i * M:(i + 1) * M]
                p = pperm[i * M:(i + 1) * M]
                idx = indices[i * M:(i + 1) * M]
                if self.torch_rng:
                    torch.randn(M, A, device=indices.device)     # the reference's unused sample() (ppo.py:154) burns RNG
                inputs = {"actor": x, "critic": p, "base_lin_vel": x}
                cols = {"actor": k_obs, "critic": k_priv, "base_lin_vel": k_obs}
                eng.forward_hidden(nets, inputs, cols, ws, M, frames=frames)
                if self.sym_loss:
                    sym = self._sym_forward(x, ws, M, frames)
                largs = HgLossArgs(
                    actor=self._head_net(ac.net_actor, ws, flat), critic=self._head_net(ac.net_critic, ws, flat),
                    linvel=self._head_net(ac.net_linvel, ws, flat), rows=M, A=A, idx=ptr(idx),
                    actions=ptr(st.actions), old_mu=ptr(st.mu), old_sigma=ptr(st.sigma),
                    old_log_prob=ptr(st.actions_log_prob), advantages=ptr(st.advantages), returns=ptr(st.returns),
                    old_values=ptr(st.values), critic_obs=cobs_ptr, ld_cobs=ld_cobs, lv_col=lv_col,
                    std=ptr(ac.std.data), g_std=ptr(flat.grad_view(ac.std)), clip_param=float(self.clip_param),
                    value_loss_coef=float(self.value_loss_coef), entropy_coef=float(self.entropy_coef),
                    lin_vel_coef=float(self.base_lin_vel_coef),
                    use_clipped_value_loss=1 if self.use_clipped_value_loss else 0,
                    inv_world=1.0 / self.world_size, stats=ptr(flat.stats))
                check(lib.hg_ppo_head_loss(C.byref(largs), current_stream()))
                if self.sym_loss:
                    self._sym_backward(sym, ws, M, frames)
                eng.backward_hidden(nets, inputs, cols, ws, M, frames=frames)
                if self.world_size > 1:
                    # the only collective of the path: SUM over ranks of [gradients | KL and loss sums]
                    self._allreduce_grads()
                self.optimizer.fused_step(rows_global=M * self.world_size, max_grad_norm=self.max_grad_norm,
                                          adaptive=adaptive, desired_kl=self.desired_kl, loss_log=loss_log, log_slot=slot)
                slot += 1
                if _lib.NVTX:
                    torch.cuda.nvtx.range_pop()
        self.last_update_enqueue_ms = (time.perf_counter() - t_host0) * 1e3    # host time to enqueue the whole update
        log = loss_log.cpu()                                      # the single host sync of update()
        self.last_update_log = log
        self.learning_rate = float(log[-1, 5])
        self._lr_synced = self.learning_rate
        for g in self.optimizer.param_groups:
            g['lr'] = self.learning_rate
        mean_value_loss = float(log[:, 0].mean())
        mean_surrogate_loss = float(log[:, 1].mean())
        mean_base_lin_vel_loss = float(log[:, 2].mean())
        st.clear()
        sym_loss = 0
        if self.sym_loss:      # the reference returns the LAST minibatch's symmetry loss (ppo.py:226)
            self.last_sym_losses = (self._sym_sums / (M * A)).cpu().tolist()
            sym_loss = self.last_sym_loss = self.last_sym_losses[-1]
        return mean_value_loss, mean_surrogate_loss, sym_loss, mean_base_lin_vel_loss

    # ------------------------------------------------------------------ mirror-symmetry loss (ppo.py:197-202)
    def _sym_perm(self, frames):
        """(source column, sign) of the mirrored observation row, in the column layout of the minibatch rows: packed
        (f*width + c) for the stacked format, frame-pitched (f*pitch + c, pad columns -> 0) for the history ring."""
        if frames is None:
            return self._obs_src, self._obs_sign, self.storage.obs_cols
        width, pitch = frames["actor"]
        key = (width, pitch)
        if getattr(self, "_sym_pitched", None) is None or self._sym_pitched[0] != key:
            src, sign = self._obs_src.cpu().tolist(), self._obs_sign.cpu().tolist()
            nf = len(src) // width
            psrc, psign = list(range(nf * pitch)), [0.0] * (nf * pitch)
            for j, (sj, gj) in enumerate(zip(src, sign)):
                f, c = divmod(j, width)
                fs, cs = divmod(sj, width)
                psrc[f * pitch + c], psign[f * pitch + c] = fs * pitch + cs, gj
            dev = self._obs_src.device
            self._sym_pitched = (key, torch.tensor(psrc, dtype=torch.int32, device=dev), torch.tensor(psign, dtype=torch.float32, device=dev), nf * pitch)
        return self._sym_pitched[1:]

    def _sym_forward(self, x, ws, M, frames=None):
        """mirror_obs = obs @ obs_perm_mat, mirror_act = actor(mirror_obs), sym_loss = mean((mu - mirror_act @ act_perm_mat)^2):
        the mirrored rows go through the same grouped tcgen05 GEMMs as the batch; returns what `_sym_backward` needs."""
        from .engine import Workspace
        ac, st, lib = self.actor_critic, self.storage, self._lib
        eng, A, dev = ac.engine(), ac.num_actions, x.device
        obs_src, obs_sign, k_obs = self._sym_perm(frames)
        if self._sym_ws is None or self._sym_ws[0].rows != M or self._sym_ws[1].shape[1] != x.shape[1]:
            f32 = lambda *s: torch.empty(*s, device=dev)
            self._sym_ws = (Workspace((ac.net_actor,), M, dev, True), torch.zeros(M, x.shape[1], dtype=torch.bfloat16, device=dev),
                            f32(M, A), f32(M, A), f32(M, A), f32(M, A))
        ws_m, xm, mu, mirror, dmu, dmirror = self._sym_ws
        if getattr(self, "_sym_slot", None) is None or self._sym_slot >= self._sym_sums.numel():
            self._sym_sums = torch.zeros(self.num_learning_epochs * self.num_mini_batches, dtype=torch.float64, device=dev)
            self._sym_slot = 0
        s = current_stream()
        check(lib.hg_permute_cols_bf16(ptr(x), x.stride(0), M, k_obs, ptr(obs_src), ptr(obs_sign), ptr(xm), xm.stride(0), s))
        afr = {"actor": frames["actor"]} if frames else None
        eng.forward_hidden((ac.net_actor,), {"actor": xm}, {"actor": k_obs}, ws_m, M, frames=afr)
        la = ac.net_actor.linears[-1]
        for h, out in ((ws.h["actor"][-1], mu), (ws_m.h["actor"][-1], mirror)):
            check(lib.hg_head_forward(ptr(h), h.stride(0), h.shape[1], ptr(la.weight.data), ptr(la.bias.data), A, M, ptr(out), s))
        scale = 2.0 * float(self.sym_coef) / (M * A * self.world_size)
        check(lib.hg_sym_loss(ptr(mu), ptr(mirror), ptr(self._act_src), ptr(self._act_sign), M, A, scale, ptr(dmu), ptr(dmirror),
                              self._sym_sums.data_ptr() + 8 * self._sym_slot, s))
        self._sym_slot += 1
        return ws_m, xm, dmu, dmirror

    def _sym_backward(self, sym, ws, M, frames=None):
        """Adds the symmetry term's gradient: through the batch's own forward (d/d mu, accumulated onto the dZ the loss
        kernel just wrote) and through the mirrored forward (its own backward pass over the actor)."""
        ws_m, xm, dmu, dmirror = sym
        ac, st, lib, flat = self.actor_critic, self.storage, self._lib, self.actor_critic.flat
        la, top = ac.net_actor.linears[-1], ac.net_actor.linears[-2]
        gW, gb, gbh = flat.grad_view(la.weight), flat.grad_view(la.bias), flat.grad_view(top.bias)
        s = current_stream()
        for h, d, dz, acc in ((ws.h["actor"][-1], dmu, ws.dz["actor"][-1], 1), (ws_m.h["actor"][-1], dmirror, ws_m.dz["actor"][-1], 0)):
            check(lib.hg_head_backward(ptr(h), h.stride(0), h.shape[1], ptr(la.weight.data), ac.num_actions, ptr(d), M, acc,
                                       ptr(dz), dz.stride(0), ptr(gW), ptr(gb), ptr(gbh), s))
        k_obs = self._sym_perm(frames)[2]
        ac.engine().backward_hidden((ac.net_actor,), {"actor": xm}, {"actor": k_obs}, ws_m, M,
                                    frames={"actor": frames["actor"]} if frames else None)
