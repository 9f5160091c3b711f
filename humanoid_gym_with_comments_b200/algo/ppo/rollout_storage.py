"""`RolloutStorage` with the reference's public tensors and methods
(reference: humanoid/algo/ppo/rollout_storage.py:35-191), GAE on a CUDA scan kernel.

Public tensors keep the reference's names, dtypes and [T, N, ...] shapes.  Observations live in one of two formats:

* stacked (generic API: any `act(obs, critic_obs)` / `add_transitions`): the reference's fp32 `[T, N, O]` / `[T, N, P]`
  tensors + bf16, 128-byte-row-aligned shadows (`obs16`, `priv16`: [T*N, round_up(width, 64)]) which are what the
  tensor-core layers read; both are written by the ingest kernel.  Allocated on first use.
* history ring (enabled by `SyntheticXBotEnv.bind_storage` / `enable_ring`): the env writes ONE new frame
  per step into env-major rings `[N, T+F, So]`; the stacked observation of row t is the window `[n, t .. t+F-1, :]`.
  `observations` / `privileged_observations` are then materialised on access (a kernel re-stacks the rows from the
  archive ring), bit-identical to what the stacked format holds.  15x less observation memory and ~6x less env traffic.
"""
from __future__ import annotations

import torch

import ctypes as C

from ... import _lib
from ..._lib import HgHistoryRing, check, current_stream, ptr
from .engine import LD_ALIGN, round_up

ALL_VALID = -(1 << 24)      # "first valid frame time" of an env that has not been reset within the ring's memory


class _Ring:
    """One history ring (actor or critic observations): tensors + the C struct handed to the kernels."""

    def __init__(self, N, rows, width, frames, device):
        self.N, self.rows, self.width, self.frames = N, rows, width, frames
        self.pitch16 = round_up(width, 8)
        self.slots = rows + frames - 1
        z = lambda *s, **k: torch.zeros(*s, device=device, **k)
        self.live32 = z(N, self.slots, width)
        self.arch32 = z(N, self.slots, width)
        self.live16 = z(N, self.slots, self.pitch16, dtype=torch.bfloat16)
        self.arch16 = z(N, self.slots, self.pitch16, dtype=torch.bfloat16)
        self.c = HgHistoryRing(live32=ptr(self.live32), live16=ptr(self.live16), arch32=ptr(self.arch32), arch16=ptr(self.arch16), width=width,
                               pitch16=self.pitch16, frames=frames, slots=self.slots)
        self.K16 = frames * self.pitch16               # bf16 window width = K of the first layer
        self._flat32 = self.live32.view(N, self.slots * width)
        self._flat16 = self.live16.view(N, self.slots * self.pitch16)
        self._v32, self._v16 = {}, {}

    def view32(self, row):
        """fp32 [N, frames*width] strided view: the stacked observation of storage row `row` (the reference's obs_buf)."""
        v = self._v32.get(row)
        if v is None:
            v = self._v32[row] = self._flat32[:, row * self.width:(row + self.frames) * self.width]
        return v

    def view16(self, row):
        """bf16 [N, frames*pitch16] strided view of the same window: the first layer's A operand (TMA reads it in place)."""
        v = self._v16.get(row)
        if v is None:
            v = self._v16[row] = self._flat16[:, row * self.pitch16:(row + self.frames) * self.pitch16]
        return v

    def ptr32(self, row):
        return self.live32.data_ptr() + 4 * row * self.width


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

        # Core.  Observation buffers are created on first use (`_ensure_stacked` / `enable_ring`).
        self._has_priv = privileged_obs_shape[0] is not None
        self._obs_all = self._priv_all = self._obs16_all = self._priv16_all = None
        self._ring_obs = self._ring_priv = None
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

        # bf16 shadows for the tensor-core layers (stacked format)
        self.obs_cols = int(obs_shape[0])
        self.obs_ld = round_up(self.obs_cols, LD_ALIGN)
        if self._has_priv:
            self.priv_cols = int(privileged_obs_shape[0])
            self.priv_ld = round_up(self.priv_cols, LD_ALIGN)
        else:
            self.priv_cols, self.priv_ld = self.obs_cols, self.obs_ld
        self._gae_scratch = z(4, dtype=torch.float64)
        self._slots = {}
        self._ring_version = 0          # bumped by everything that writes the rings (env step, rollover)
        self._mat = {}                  # materialised [T, N, O] / [T, N, P] views of the rings: name -> (version, tensor)
        self._rolled_from = 0

    # ------------------------------------------------------------------ observation formats
    def _ensure_stacked(self):
        """Allocate the stacked fp32 tensors + bf16 shadows.  One extra slot (index T) behind the public [T, N, ...]
        views: an env bound without a ring writes the observation of step t+1 straight into slot t+1, so slot T carries
        the last observation of an iteration over to the next one."""
        if self._obs_all is not None:
            return
        T, N = self.num_transitions_per_env, self.num_envs
        z = lambda *s, **k: torch.zeros(*s, device=self.device, **k)
        self._obs_all = z(T + 1, N, *self.obs_shape)
        self._obs16_all = z((T + 1) * N, self.obs_ld, dtype=torch.bfloat16)
        if self._has_priv:
            self._priv_all = z(T + 1, N, *self.privileged_obs_shape)
            self._priv16_all = z((T + 1) * N, self.priv_ld, dtype=torch.bfloat16)
        else:
            self._priv16_all = self._obs16_all

    @property
    def ring_enabled(self) -> bool:
        return self._ring_obs is not None

    def enable_ring(self, single_obs: int, frames: int, single_priv: int, c_frames: int):
        """Switch the observation format to the history ring (called by an env that writes single frames).  From here on
        `PPO.act` must be handed the ring's own windows (what the env returns); `disable_ring()` goes back."""
        if not self._has_priv:
            raise _lib.HgError("the history ring needs privileged (critic) observations")
        if single_obs * frames != self.obs_cols or single_priv * c_frames != self.priv_cols:
            raise _lib.HgError(f"history ring: {frames} x {single_obs} / {c_frames} x {single_priv} frames do not match the storage's "
                               f"observation widths {self.obs_cols} / {self.priv_cols}")
        if self.ring_enabled:
            return
        T, N = self.num_transitions_per_env, self.num_envs
        self._ring_obs = _Ring(N, T + 1, single_obs, frames, self.device)
        self._ring_priv = _Ring(N, T + 1, single_priv, c_frames, self.device)
        self._ring_pair = (HgHistoryRing * 2)(self._ring_obs.c, self._ring_priv.c)
        self.ring_first_valid = torch.full((T + 1, N), ALL_VALID, dtype=torch.int32, device=self.device)
        self.ring_hist_start = torch.full((N,), ALL_VALID, dtype=torch.int32, device=self.device)
        self._ring_version += 1
        self._rolled_from = 0

    def disable_ring(self):
        self._ring_obs = self._ring_priv = None
        self._mat.clear()

    def ring_import(self, obs: torch.Tensor, priv: torch.Tensor, row: int = 0):
        """Load stacked observations [N, F*So] / [N, Fc*Sp] (e.g. the env's current obs_buf) as the window of `row`."""
        with torch.no_grad():
            for ring, x in ((self._ring_obs, obs), (self._ring_priv, priv)):
                fr = x.reshape(self.num_envs, ring.frames, ring.width).float()
                for t in (ring.live32, ring.arch32):
                    t[:, row:row + ring.frames].copy_(fr)
                ring.live16[:, row:row + ring.frames, :ring.width].copy_(fr)
                ring.arch16[:, row:row + ring.frames, :ring.width].copy_(fr)
            self.ring_hist_start.fill_(ALL_VALID)
            self.ring_first_valid[row].fill_(ALL_VALID)
        self._ring_version += 1

    def ring_rollover(self, from_row: int):
        """The window of `from_row` becomes row 0 (start of the next iteration)."""
        if from_row > 0:
            check(_lib.load().hg_ring_rollover(self._ring_pair, 2, self.num_envs, int(from_row), ptr(self.ring_hist_start),
                                               ptr(self.ring_first_valid), current_stream()))
            self._ring_version += 1
        self._rolled_from = from_row

    def _materialise(self, name: str):
        """[T, N, F*width] fp32 out of the archive ring (frames older than a row's last reset are zero): what the stacked
        format would hold, bit for bit."""
        ring = self._ring_obs if name == "obs" else self._ring_priv
        hit = self._mat.get(name)
        if hit is not None and hit[0] == self._ring_version:
            return hit[1]
        T, N = self.num_transitions_per_env, self.num_envs
        out = hit[1] if hit is not None else torch.empty(T, N, ring.frames * ring.width, device=self.device)
        check(_lib.load().hg_ring_gather(C.byref(ring.c), N, ptr(self.ring_first_valid), None, T * N, ptr(out), out.shape[-1], 0,
                                         None, 0, current_stream()))
        self._mat[name] = (self._ring_version, out)
        return out

    @property
    def observations(self):
        if self.ring_enabled:
            return self._materialise("obs")
        self._ensure_stacked()
        return self._obs_all[:self.num_transitions_per_env]

    @property
    def privileged_observations(self):
        if not self._has_priv:
            return None
        if self.ring_enabled:
            return self._materialise("priv")
        self._ensure_stacked()
        return self._priv_all[:self.num_transitions_per_env]

    @property
    def obs16(self):
        self._ensure_stacked()
        return self._obs16_all[:self.num_transitions_per_env * self.num_envs]

    @property
    def priv16(self):
        self._ensure_stacked()
        return self._priv16_all[:self.num_transitions_per_env * self.num_envs]

    # ------------------------------------------------------------------ fused-path helpers (used by PPO)
    def check_room(self):
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")

    def slot(self, step: int):
        """(obs fp32 [N,O], priv fp32 [N,P] or None, obs bf16 [N,ldO], priv bf16 [N,ldP]) of slot `step` in [0, T]."""
        cached = self._slots.get(step)
        if cached is None:
            self._ensure_stacked()
            N = self.num_envs
            cached = (self._obs_all[step], self._priv_all[step] if self._priv_all is not None else None,
                      self._obs16_all[step * N:(step + 1) * N], self._priv16_all[step * N:(step + 1) * N])
            self._slots[step] = cached
        return cached

    def holds(self, obs: torch.Tensor, critic_obs, step: int) -> bool:
        """True if `obs` / `critic_obs` ARE slot `step` (written there, with their bf16 shadows, by a bound env)."""
        if self.ring_enabled:
            want = self._ring_obs.ptr32(step)
            # row 0 is also known by the address it had as the LAST row of the previous iteration (what the runner still
            # holds after `clear()` moved that window to the front of the ring)
            return obs.data_ptr() == want or (step == 0 and obs.data_ptr() == self._ring_obs.ptr32(self._rolled_from))
        if self._obs_all is None:
            return False
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
        if self.ring_enabled:
            raise _lib.HgError("this storage is bound to an env's history ring: hand PPO.act the observations that env returns, "
                               "or unbind it (env.bind_storage(None)) to store foreign observations")
        obs = obs if obs.stride(-1) == 1 else obs.contiguous()
        o16 = self.obs16[step * N:(step + 1) * N]
        dst = self.observations[step]
        same = obs.data_ptr() == dst.data_ptr()
        check(lib.hg_ingest_rows(ptr(obs), obs.stride(0), N, self.obs_cols, 0 if same else ptr(dst), dst.stride(0),
                                 ptr(o16), self.obs_ld, current_stream()))
        if self._has_priv:
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
        if self.ring_enabled:
            self.ring_rollover(self.step)      # the last row written carries over: it is row 0 of the next iteration
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
        if self._has_priv:
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
