"""`ActorCritic` with the reference's constructor, sub-modules, state_dict keys and methods
(reference: humanoid/algo/ppo/actor_critic.py:36-149), computing on sm_100a kernels.

Layout kept for drop-in use: `.actor`, `.base_lin_vel`, `.critic` are `nn.Sequential` of `nn.Linear` / `nn.ELU`
(so `state_dict()` keys are `std`, `actor.{0,2,4,6}.*`, `base_lin_vel.{0,2,4}.*`, `critic.{0,2,4,6}.*` and
`export_policy_as_jit` can deepcopy/script `.actor`), `std` is an `nn.Parameter [A]`.  On CUDA the parameters
become views of one flat arena (see engine.FlatParams); hidden layers run as grouped tcgen05 GEMMs on bf16
shadows, the output layers in fp32.  There is no CPU path: calling a compute method with CPU tensors raises.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from ... import _lib
from ..._lib import check, current_stream, ptr
from .engine import LD_ALIGN, FlatParams, MLPEngine, NetSpec, Workspace, round_up


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions,
                 actor_hidden_dims=[256, 256, 256], critic_hidden_dims=[256, 256, 256],
                 base_lin_vel_hidden_dims=[128, 128], init_noise_std=1.0, activation=nn.ELU(), **kwargs):
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: "
                  + str([key for key in kwargs.keys()]))
        super().__init__()
        if isinstance(activation, str):
            if activation.lower() != "elu":
                raise NotImplementedError("only ELU has a fused kernel")
            activation = nn.ELU()

        def mlp(n_in, hidden, n_out):
            layers, prev = [], n_in
            for h in hidden:
                layers += [nn.Linear(prev, h), activation]
                prev = h
            layers.append(nn.Linear(prev, n_out))
            return nn.Sequential(*layers)

        # construction order = RNG consumption order of the reference ctor: actor, base_lin_vel, critic
        self.actor = mlp(num_actor_obs, actor_hidden_dims, num_actions)
        self.base_lin_vel = mlp(num_actor_obs, base_lin_vel_hidden_dims, 3)
        self.critic = mlp(num_critic_obs, critic_hidden_dims, 1)
        print(f"Actor MLP: {self.actor}")
        print(f"Lin vel MLP: {self.base_lin_vel}")
        print(f"Critic MLP: {self.critic}")
        self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        self.distribution = None
        self.num_actor_obs, self.num_critic_obs, self.num_actions = num_actor_obs, num_critic_obs, num_actions
        # engine state (built lazily on the first CUDA use)
        self._flat: Optional[FlatParams] = None
        self._engine: Optional[MLPEngine] = None
        self._ws: Dict[int, Workspace] = {}
        self._x16: Dict[tuple, torch.Tensor] = {}
        self._mean = None
        self._rng_counter = 0
        self._load_version = 0

    # ------------------------------------------------------------------ engine plumbing
    def _specs(self):
        return (NetSpec("actor", self.actor), NetSpec("critic", self.critic), NetSpec("base_lin_vel", self.base_lin_vel))

    def engine(self) -> MLPEngine:
        if self._flat is None or not self._flat.attached():
            if not self.std.is_cuda:
                raise _lib.HgError("ActorCritic must be moved to a CUDA device before use (no CPU fallback)")
            self.net_actor, self.net_critic, self.net_linvel = self._specs()
            self._flat = FlatParams(self, (self.net_actor, self.net_critic, self.net_linvel))
            self._engine = MLPEngine(self._flat)
            self._ws.clear()
            self._x16.clear()
        return self._engine

    @property
    def flat(self) -> FlatParams:
        self.engine()
        return self._flat

    @staticmethod
    def philox_seed(rank: int = 0) -> int:
        """64-bit key of the in-kernel Philox streams: torch's seed, mixed with the data-parallel rank."""
        return (int(torch.initial_seed()) ^ (rank * 0x9E3779B97F4A7C15)) & (2 ** 64 - 1)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._load_version += 1
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        if self._flat is not None and self._flat.attached():
            self._flat.refresh_shadows()
        return out

    def workspace(self, rows: int, with_grads: bool = False) -> Workspace:
        key = rows if not with_grads else -rows
        ws = self._ws.get(key)
        if ws is None:
            nets = (self.net_actor, self.net_critic, self.net_linvel)
            ws = Workspace(nets, rows, self.std.device, with_grads)
            self._ws[key] = ws
        return ws

    def _to_bf16(self, x: torch.Tensor, tag: str) -> torch.Tensor:
        """fp32 [rows, cols] -> cached bf16 [rows, round_up(cols,64)] through the ingest kernel."""
        _lib.require_cuda(x)
        if x.dtype != torch.float32 or x.dim() != 2:
            raise _lib.HgError("observations must be float32 [rows, features]")
        if x.stride(1) != 1:
            x = x.contiguous()
        rows, cols = x.shape
        ld = round_up(cols, LD_ALIGN)
        buf = self._x16.get((tag, rows, ld))
        if buf is None:
            buf = torch.empty(rows, ld, dtype=torch.bfloat16, device=x.device)
            self._x16[(tag, rows, ld)] = buf
        check(_lib.load().hg_ingest_rows(ptr(x), x.stride(0), rows, cols, 0, 0, ptr(buf), ld, current_stream()))
        return buf

    def _head_rollout(self, ws, rows, with_actor, eps, out):
        lib = _lib.load()
        A = self.num_actions
        la, lc = self.net_actor.linears[-1], self.net_critic.linears[-1]
        ha, hc = ws.h["actor"][-1], ws.h["critic"][-1]
        self._rng_counter += 1
        if with_actor:
            check(lib.hg_policy_head_rollout(
                ptr(ha), ha.stride(0), ha.shape[1], ptr(la.weight.data), ptr(la.bias.data), ptr(self.std.data),
                ptr(hc), hc.stride(0), hc.shape[1], ptr(lc.weight.data), ptr(lc.bias.data), ptr(eps),
                self.philox_seed(torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0),
                self._rng_counter, rows, A,
                ptr(out["actions"]), ptr(out["mu"]), ptr(out["sigma"]), ptr(out["log_prob"]), ptr(out["values"]),
                current_stream()))
        else:
            check(lib.hg_value_head(ptr(hc), hc.stride(0), hc.shape[1], ptr(lc.weight.data), ptr(lc.bias.data), rows,
                                    ptr(out["values"]), current_stream()))

    # ------------------------------------------------------------------ reference API
    @staticmethod
    def init_weights(sequential, scales):
        [torch.nn.init.orthogonal_(module.weight, gain=scales[idx]) for idx, module in
         enumerate(mod for mod in sequential if isinstance(mod, nn.Linear))]

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self._mean

    @property
    def action_std(self):
        return self._mean * 0.0 + self.std

    @property
    def entropy(self):
        # Normal.entropy().sum(-1): depends on std only (actor_critic.py:123-125)
        e = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.std)).sum()
        return e.expand(self._mean.shape[0])

    def _net_forward(self, net: NetSpec, x: torch.Tensor) -> torch.Tensor:
        """Full forward of one net for arbitrary rows: hidden layers on tensor cores, output layer through
        fp32 head kernel."""
        eng = self.engine()
        rows = x.shape[0]
        x16 = self._to_bf16(x, net.name)
        ws = self.workspace(rows)
        eng.forward_hidden((net,), {net.name: x16}, {net.name: x.shape[1]}, ws, rows)
        h = ws.h[net.name][-1]
        last = net.linears[-1]
        out = torch.empty(rows, last.out_features, device=x.device)
        check(_lib.load().hg_head_forward(ptr(h), h.stride(0), h.shape[1], ptr(last.weight.data), ptr(last.bias.data),
                                          last.out_features, rows, ptr(out), current_stream()))
        return out

    def update_distribution(self, observations):
        self.engine()
        self._mean = self._net_forward(self.net_actor, observations)

    @property
    def distribution(self):
        """The reference's `self.distribution` attribute (a torch Normal over the last means), built on access only: `act`
        and `get_actions_log_prob` run on the library's kernels."""
        if self._mean is None:
            return None
        return torch.distributions.Normal(self._mean, self._mean * 0.0 + self.std.data, validate_args=False)

    @distribution.setter
    def distribution(self, value):
        pass

    def act(self, observations, **kwargs):
        self.engine()
        self.update_distribution(observations)
        rows, A = self._mean.shape
        # the reference draws Normal(mean, std).sample() from torch's generator: the same stream as randn of that shape
        eps = torch.randn(rows, A, device=self._mean.device)
        action = torch.empty_like(self._mean)
        self._rng_counter += 1
        check(_lib.load().hg_gaussian_sample(ptr(self._mean), ptr(self.std.data), ptr(eps), 0, self._rng_counter, rows, A, ptr(action), None,
                                             current_stream()))
        base_lin_vel = self.base_get_lin_vel(observations)
        return action, base_lin_vel

    def get_actions_log_prob(self, actions):
        _lib.require_cuda(actions)
        a = actions if (actions.dtype == torch.float32 and actions.is_contiguous()) else actions.float().contiguous()
        rows, A = self._mean.shape
        out = torch.empty(rows, device=self._mean.device)
        check(_lib.load().hg_gaussian_log_prob(ptr(a), ptr(self._mean), ptr(self.std.data), rows, A, ptr(out), current_stream()))
        return out

    def act_inference(self, observations):
        self.engine()
        return self._net_forward(self.net_actor, observations)

    def evaluate(self, critic_observations, **kwargs):
        self.engine()
        return self._net_forward(self.net_critic, critic_observations)

    def base_get_lin_vel(self, observations):
        self.engine()
        return self._net_forward(self.net_linvel, observations)
