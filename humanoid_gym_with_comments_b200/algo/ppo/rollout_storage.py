"""`RolloutStorage` with the reference's public tensors and methods
(reference: humanoid/algo/ppo/rollout_storage.py:35-191), GAE on a CUDA scan kernel.

Public tensors keep the reference's names, dtypes and [T, N, ...] shapes.  In addition the storage owns bf16,
128-byte-row-aligned shadows of the observations (`obs16`, `priv16`: [T*N, round_up(width, 64)]) which are what
the tensor-core layers read; they are written by the same ingest kernel that fills the fp32 slot.
"""
from __future__ import annotations

import torch

from ... import _lib
from ..._lib import check, current_stream, ptr
from .engine import LD_ALIGN, round_up


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device='cpu'):
        self.device = device
        if not torch.device(device).type == "cuda":
            raise _lib.HgError("RolloutStorage lives on a CUDA device (no CPU fallback); got device=" + str(device))
        self.obs_shape = obs_shape
        self.privileged_obs_shape = privileged_obs_shape
        self.actions_shape = actions_shape
        T, N = num_transitions_per_env, num_envs
        z = lambda *s, **k: torch.zeros(*s, device=self.device, **k)

        # Core.  One extra slot (index T) behind the public [T, N, ...] views: an env bound with `bind_env` writes
        # the observation of step t+1 straight into slot t+1, so slot T carries the last observation of an
        # iteration over to the next one.
        self._obs_all = z(T + 1, N, *obs_shape)
        self.observations = self._obs_all[:T]
        if privileged_obs_shape[0] is not None:
            self._priv_all = z(T + 1, N, *privileged_obs_shape)
            self.privileged_observations = self._priv_all[:T]
        else:
            self._priv_all = None
            self.privileged_observations = None
        self.rewards = z(T, N, 1)
        self.actions = z(T, N, *actions_shape)
        self.dones = z(T, N, 1).byte()

        # For PPO
        self.actions_log_prob = z(T, N, 1)
        self.values = z(T, N, 1)
        self.returns = z(T, N, 1)
        self.advantages = z(T, N, 1)
        self.mu = z(T, N, *actions_shape)
        self.sigma = z(T, N, *actions_shape)

        self.num_transitions_per_env = T
        self.num_envs = N

        # rnn (unused by the reference's runner; kept for attribute compatibility)
        self.saved_hidden_states_a = None
        self.saved_hidden_states_c = None

        self.step = 0

        # bf16 shadows for the tensor-core layers
        self.obs_cols = int(obs_shape[0])
        self.obs_ld = round_up(self.obs_cols, LD_ALIGN)
        self._obs16_all = z((T + 1) * N, self.obs_ld, dtype=torch.bfloat16)
        self.obs16 = self._obs16_all[:T * N]
        if self.privileged_observations is not None:
            self.priv_cols = int(privileged_obs_shape[0])
            self.priv_ld = round_up(self.priv_cols, LD_ALIGN)
            self._priv16_all = z((T + 1) * N, self.priv_ld, dtype=torch.bfloat16)
            self.priv16 = self._priv16_all[:T * N]
        else:
            self.priv_cols, self.priv_ld = self.obs_cols, self.obs_ld
            self._priv16_all, self.priv16 = self._obs16_all, self.obs16
        self._gae_scratch = z(4, dtype=torch.float64)
        self._slots = {}

    # ------------------------------------------------------------------ fused-path helpers (used by PPO)
    def check_room(self):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")

    def slot(self, step: int):
        """(obs fp32 [N,O], priv fp32 [N,P] or None, obs bf16 [N,ldO], priv bf16 [N,ldP]) of slot `step` in [0, T]."""
        cached = self._slots.get(step)
        if cached is None:
            N = self.num_envs
            cached = (self._obs_all[step], self._priv_all[step] if self._priv_all is not None else None,
                      self._obs16_all[step * N:(step + 1) * N], self._priv16_all[step * N:(step + 1) * N])
            self._slots[step] = cached
        return cached

    def holds(self, obs: torch.Tensor, critic_obs, step: int) -> bool:
        """True if `obs` / `critic_obs` ARE slot `step` (written there, with their bf16 shadows, by a bound env)."""
        o, p, _, _ = self.slot(step)
        if obs is not o and obs.data_ptr() != o.data_ptr():
            return False
        return p is None or critic_obs is None or critic_obs is p or critic_obs.data_ptr() == p.data_ptr()

    def ingest(self, obs: torch.Tensor, critic_obs, step: int):
        """Copy this step's observations into slot `step` (fp32) and its bf16 shadow rows, one launch each
        (replaces the two large copy_ of add_transitions, rollout_storage.py:90-91).
        Returns the bf16 row blocks (obs16_t, priv16_t)."""
        lib = _lib.load()
        N = self.num_envs
        _lib.require_cuda(obs, critic_obs)
        obs = obs if obs.stride(-1) == 1 else obs.contiguous()
        o16 = self.obs16[step * N:(step + 1) * N]
        dst = self.observations[step]
        same = obs.data_ptr() == dst.data_ptr()
        check(lib.hg_ingest_rows(ptr(obs), obs.stride(0), N, self.obs_cols, 0 if same else ptr(dst), dst.stride(0),
                                 ptr(o16), self.obs_ld, current_stream()))
        if self.privileged_observations is not None:
            c = critic_obs if critic_obs.stride(-1) == 1 else critic_obs.contiguous()
            p16 = self.priv16[step * N:(step + 1) * N]
            dstp = self.privileged_observations[step]
            samep = c.data_ptr() == dstp.data_ptr()
            check(lib.hg_ingest_rows(ptr(c), c.stride(0), N, self.priv_cols, 0 if samep else ptr(dstp),
                                     dstp.stride(0), ptr(p16), self.priv_ld, current_stream()))
        else:
            p16 = o16
        return o16, p16

    # ------------------------------------------------------------------ reference API
    def add_transitions(self, transition: Transition):
        """Generic (non-fused) path with the reference's semantics, for external callers."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s = self.step
        self.ingest(transition.observations, transition.critic_observations, s)
        self.actions[s].copy_(transition.actions)
        self.rewards[s].copy_(transition.rewards.view(-1, 1))
        self.dones[s].copy_(transition.dones.view(-1, 1))
        self.values[s].copy_(transition.values)
        self.actions_log_prob[s].copy_(transition.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(transition.action_mean)
        self.sigma[s].copy_(transition.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """One reverse-time scan kernel + one normalisation kernel (rollout_storage.py:122-143)."""
        _lib.require_cuda(last_values)
        lv = last_values.reshape(-1).contiguous().float()
        T, N = self.num_transitions_per_env, self.num_envs
        check(_lib.load().hg_gae_returns(ptr(self.rewards), ptr(self.values), ptr(self.dones), ptr(lv),
                                         float(gamma), float(lam), T, N, ptr(self.returns), ptr(self.advantages),
                                         ptr(self._gae_scratch), current_stream()))

    def get_statistics(self):
        done = self.dones
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = (done_indices[1:] - done_indices[:-1])
        return trajectory_lengths.float().mean(), self.rewards.mean()

    def draw_permutation(self, num_mini_batches):
        """`torch.randperm` on the storage device, exactly as the reference draws it (rollout_storage.py:154-156):
        once per update, reused by every epoch."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = torch.randperm(num_mini_batches * mini_batch_size, requires_grad=False, device=self.device)
        return indices, mini_batch_size

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """Reference-shaped generator (12-tuple, rollout_storage.py:153-191) for external callers.  `PPO.update`
        does not use it: it consumes the same permutation through fused gather kernels instead."""
        indices, mini_batch_size = self.draw_permutation(num_mini_batches)
        observations = self.observations.flatten(0, 1)
        if self.privileged_observations is not None:
            critic_observations = self.privileged_observations.flatten(0, 1)
        else:
            critic_observations = observations
        base_lin_vel = critic_observations[:, 53:56]
        actions = self.actions.flatten(0, 1)
        values = self.values.flatten(0, 1)
        returns = self.returns.flatten(0, 1)
        old_actions_log_prob = self.actions_log_prob.flatten(0, 1)
        advantages = self.advantages.flatten(0, 1)
        old_mu = self.mu.flatten(0, 1)
        old_sigma = self.sigma.flatten(0, 1)
        for epoch in range(num_epochs):
            for i in range(num_mini_batches):
                batch_idx = indices[i * mini_batch_size:(i + 1) * mini_batch_size]
                yield (observations[batch_idx], critic_observations[batch_idx], base_lin_vel[batch_idx],
                       actions[batch_idx], values[batch_idx], advantages[batch_idx], returns[batch_idx],
                       old_actions_log_prob[batch_idx], old_mu[batch_idx], old_sigma[batch_idx], (None, None), None)
