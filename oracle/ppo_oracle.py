"""CPU oracle (torch fp32) for the learner half of the hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package.

A functional restatement of the reference learner, pinned against the unmodified
reference by `oracle/make_golden.py` (fixtures under tests/golden/):

  reference file:line                              -> here
  humanoid/algo/ppo/actor_critic.py:37-99  (ctor)  -> init_params
  actor_critic.py:127-138 (Normal, sample, logp)   -> policy_mean / sample_actions / log_prob / entropy
  actor_critic.py:144-148 (critic, lin-vel head)   -> value / lin_vel
  humanoid/algo/ppo/ppo.py:128-138                 -> bootstrap_rewards
  rollout_storage.py:122-143 (GAE + normalise)     -> gae_returns
  rollout_storage.py:153-191 (minibatches)         -> minibatch_slices / gather_minibatch
  ppo.py:144-226 (update)                          -> minibatch_losses / adaptive_lr /
                                                      clip_grad_norm / adam_step / ppo_update

Gradients come from torch autograd, as in the reference.  Everything is float32 on CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- model
def _linear_init(out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for
    the weight, same bound for the bias), drawing from the global generator in the same order."""
    bound = 1.0 / math.sqrt(in_f)
    w = torch.empty(out_f, in_f).uniform_(-bound, bound)
    b = torch.empty(out_f).uniform_(-bound, bound)
    return w, b


def mlp_dims(n_in: int, hidden: Sequence[int], n_out: int) -> List[Tuple[int, int]]:
    dims = [n_in] + list(hidden) + [n_out]
    return [(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]


def init_params(num_actor_obs: int, num_critic_obs: int, num_actions: int,
                actor_hidden=(256, 256, 256), critic_hidden=(256, 256, 256),
                lin_vel_hidden=(128, 128), init_noise_std: float = 1.0) -> Params:
    """Same tensors, names and RNG consumption order as `ActorCritic.__init__`
    (actor_critic.py:37-99): actor, then base_lin_vel, then critic; `std` last (no RNG)."""
    p: Params = {}
    for prefix, dims in (("actor", mlp_dims(num_actor_obs, actor_hidden, num_actions)),
                         ("base_lin_vel", mlp_dims(num_actor_obs, lin_vel_hidden, 3)),
                         ("critic", mlp_dims(num_critic_obs, critic_hidden, 1))):
        for i, (fi, fo) in enumerate(dims):
            w, b = _linear_init(fo, fi)
            p[f"{prefix}.{2 * i}.weight"] = w
            p[f"{prefix}.{2 * i}.bias"] = b
    p["std"] = init_noise_std * torch.ones(num_actions)
    return p


def state_dict_order(p: Params) -> List[str]:
    """Order of `ActorCritic.parameters()` / `state_dict()` in the reference: `std` is registered
    after the three Sequentials in __init__ but nn.Module lists direct Parameters first."""
    keys = ["std"]
    for prefix in ("actor", "base_lin_vel", "critic"):
        idx = sorted({int(k.split(".")[1]) for k in p if k.startswith(prefix + ".")})
        for i in idx:
            keys += [f"{prefix}.{i}.weight", f"{prefix}.{i}.bias"]
    return keys


def mlp(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    idx = sorted({int(k.split(".")[1]) for k in p if k.startswith(prefix + ".")})
    for n, i in enumerate(idx):
        x = torch.nn.functional.linear(x, p[f"{prefix}.{i}.weight"], p[f"{prefix}.{i}.bias"])
        if n != len(idx) - 1:
            x = torch.nn.functional.elu(x)
    return x


def policy_mean(p: Params, obs: torch.Tensor) -> torch.Tensor:
    return mlp(p, "actor", obs)


def value(p: Params, critic_obs: torch.Tensor) -> torch.Tensor:
    return mlp(p, "critic", critic_obs)


def lin_vel(p: Params, obs: torch.Tensor) -> torch.Tensor:
    return mlp(p, "base_lin_vel", obs)


def sample_actions(mean: torch.Tensor, std: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """`Normal(mean, mean*0+std).sample()` == mean + std*eps with eps ~ N(0,1) drawn by one
    `normal_` of the output shape (actor_critic.py:129,133)."""
    return mean + (mean * 0.0 + std) * eps


def log_prob(mean: torch.Tensor, std: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """Sum over action dims of the Normal log-density (actor_critic.py:137-138)."""
    sigma = mean * 0.0 + std
    var = sigma ** 2
    lp = -((actions - mean) ** 2) / (2 * var) - sigma.log() - math.log(math.sqrt(2 * math.pi))
    return lp.sum(dim=-1)


def entropy(mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    sigma = mean * 0.0 + std
    return (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(sigma)).sum(dim=-1)


# ----------------------------------------------------------------------------- rollout side
def bootstrap_rewards(rewards: torch.Tensor, values: torch.Tensor, time_outs: Optional[torch.Tensor],
                      gamma: float) -> torch.Tensor:
    """ppo.py:129-133: rewards.clone() += gamma * squeeze(values * time_outs[:,None], 1)."""
    r = rewards.clone()
    if time_outs is not None:
        r += gamma * torch.squeeze(values * time_outs.unsqueeze(1), 1)
    return r


def gae_returns(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor,
                last_values: torch.Tensor, gamma: float, lam: float):
    """rollout_storage.py:122-143.  rewards/values [T,N,1] fp32, dones [T,N,1] uint8,
    last_values [N,1].  Returns (returns, normalised advantages), both [T,N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    adv = 0
    for t in reversed(range(T)):
        nxt = last_values if t == T - 1 else values[t + 1]
        not_term = 1.0 - dones[t].float()
        delta = rewards[t] + not_term * gamma * nxt - values[t]
        adv = delta + not_term * gamma * lam * adv
        returns[t] = adv + values[t]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


@dataclass
class Rollout:
    """The contents of RolloutStorage after T steps ([T,N,...], rollout_storage.py:61-76)."""
    observations: torch.Tensor
    privileged_observations: Optional[torch.Tensor]
    actions: torch.Tensor
    rewards: torch.Tensor
    dones: torch.Tensor
    values: torch.Tensor
    actions_log_prob: torch.Tensor
    mu: torch.Tensor
    sigma: torch.Tensor
    returns: Optional[torch.Tensor] = None
    advantages: Optional[torch.Tensor] = None


def minibatch_slices(perm: torch.Tensor, num_mini_batches: int, num_epochs: int):
    """rollout_storage.py:153-176: ONE permutation per update, reused every epoch;
    minibatch i takes perm[i*mb:(i+1)*mb]."""
    mb = perm.numel() // num_mini_batches
    for _ in range(num_epochs):
        for i in range(num_mini_batches):
            yield perm[i * mb:(i + 1) * mb]


LIN_VEL_SLICE = slice(53, 56)  # rollout_storage.py:164 (hard-coded in the reference)


def gather_minibatch(ro: Rollout, idx: torch.Tensor) -> Dict[str, torch.Tensor]:
    """rollout_storage.py:158-189: flatten [T,N,..] -> [T*N,..] (row = t*N+n), gather rows."""
    obs = ro.observations.flatten(0, 1)
    cobs = ro.privileged_observations.flatten(0, 1) if ro.privileged_observations is not None else obs
    return dict(
        obs=obs[idx], critic_obs=cobs[idx], lin_vel_target=cobs[:, LIN_VEL_SLICE][idx],
        actions=ro.actions.flatten(0, 1)[idx], old_values=ro.values.flatten(0, 1)[idx],
        advantages=ro.advantages.flatten(0, 1)[idx], returns=ro.returns.flatten(0, 1)[idx],
        old_log_prob=ro.actions_log_prob.flatten(0, 1)[idx],
        old_mu=ro.mu.flatten(0, 1)[idx], old_sigma=ro.sigma.flatten(0, 1)[idx])


# ----------------------------------------------------------------------------- update side
@dataclass
class PPOConfig:
    num_learning_epochs: int = 1
    num_mini_batches: int = 1
    clip_param: float = 0.2
    gamma: float = 0.998
    lam: float = 0.95
    value_loss_coef: float = 1.0
    entropy_coef: float = 0.0
    learning_rate: float = 1e-3
    max_grad_norm: float = 1.0
    use_clipped_value_loss: bool = True
    schedule: str = "fixed"
    desired_kl: Optional[float] = 0.01
    base_lin_vel_coef: float = 1.0
    # mirror-symmetry loss (ppo.py:59-63,93-105,197-202): off in the shipped config
    sym_loss: bool = False
    sym_coef: float = 1.0
    obs_permutation: Optional[List[float]] = None      # per single frame, signed indices (|p| = source column)
    act_permutation: Optional[List[float]] = None
    frame_stack: int = 0
    # torch.optim.Adam defaults (ppo.py:77)
    beta1: float = 0.9
    beta2: float = 0.999
    adam_eps: float = 1e-8


def xbot_ppo_config() -> PPOConfig:
    """humanoid_config.py:473-484 (XBotLCfgPPO.algorithm)."""
    return PPOConfig(num_learning_epochs=2, num_mini_batches=4, clip_param=0.2, gamma=0.994, lam=0.9,
                     value_loss_coef=1.0, entropy_coef=0.001, learning_rate=1e-5, max_grad_norm=1.0,
                     use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01)


def kl_old_new(mu, sigma, old_mu, old_sigma) -> torch.Tensor:
    """ppo.py:165-166, per row."""
    return torch.sum(torch.log(sigma / old_sigma + 1.e-5)
                     + (torch.square(old_sigma) + torch.square(old_mu - mu)) / (2.0 * torch.square(sigma))
                     - 0.5, dim=-1)


def adaptive_lr(lr: float, kl_mean: float, desired_kl: float) -> float:
    """ppo.py:169-172."""
    if kl_mean > desired_kl * 2.0:
        return max(1e-5, lr / 1.5)
    if kl_mean < desired_kl / 2.0 and kl_mean > 0.0:
        return min(1e-2, lr * 1.5)
    return lr


def perm_matrix(perm: List[float]) -> torch.Tensor:
    """ppo.py:96-98 / 103-105: M[int(|p_i|)][i] = sign(p_i), so (x @ M)[i] = sign(p_i) * x[|p_i|].  (An index 0 has to be
    written as +-0.0001 by the caller: np.sign(0) = 0 would zero the column -- reproduced as the reference behaves.)"""
    import numpy as np
    m = torch.zeros((len(perm), len(perm)))
    for i, q in enumerate(perm):
        m[int(abs(q))][i] = float(np.sign(q))
    return m


def stacked_obs_permutation(obs_permutation: List[float], frame_stack: int) -> List[float]:
    """ppo.py:99-102: the per-frame permutation repeated for every stacked frame, offset by the frame width."""
    import numpy as np
    out = []
    for i in range(frame_stack):
        for q in obs_permutation:
            out.append(float(np.sign(q) * (abs(q) + i * len(obs_permutation))))
    return out


def xbot_mirror_permutations():
    """A left/right mirror of the 47-wide XBot-L frame and of its 12 actions, for tests (the reference ships no
    permutation; `sym_loss` is off in humanoid_config.py).  Signs: hip roll / hip yaw / ankle roll flip."""
    s = [-1, -1, 1, 1, 1, -1]
    obs = [-0.0001, 1, 2, -3, -4]
    for b in (5, 17, 29):
        obs += [s[i] * (b + 6 + i) for i in range(6)] + [s[i] * (b + i) for i in range(6)]
    obs += [-41, 42, -43, -44, 45, -46]
    act = [s[i] * (6 + i) for i in range(6)] + [(s[i] * i if i else -0.0001) for i in range(6)]
    assert sorted(int(abs(q)) for q in obs) == list(range(47)) and sorted(int(abs(q)) for q in act) == list(range(12))
    return obs, act


def minibatch_losses(p: Params, mb: Dict[str, torch.Tensor], cfg: PPOConfig) -> Dict[str, torch.Tensor]:
    """Forward + the loss terms of ppo.py:154-208 (the mirror-symmetry term only with cfg.sym_loss)."""
    mu = policy_mean(p, mb["obs"])
    std = p["std"]
    sigma = mu * 0.0 + std
    lv = lin_vel(p, mb["obs"])
    logp = log_prob(mu, std, mb["actions"])
    v = value(p, mb["critic_obs"])
    ent = entropy(mu, std)
    with torch.no_grad():
        kl_mean = kl_old_new(mu, sigma, mb["old_mu"], mb["old_sigma"]).mean()
    adv = torch.squeeze(mb["advantages"])
    ratio = torch.exp(logp - torch.squeeze(mb["old_log_prob"]))
    surrogate = -adv * ratio
    surrogate_clipped = -adv * torch.clamp(ratio, 1.0 - cfg.clip_param, 1.0 + cfg.clip_param)
    surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
    if cfg.use_clipped_value_loss:
        v_clipped = mb["old_values"] + (v - mb["old_values"]).clamp(-cfg.clip_param, cfg.clip_param)
        value_loss = torch.max((v - mb["returns"]).pow(2), (v_clipped - mb["returns"]).pow(2)).mean()
    else:
        value_loss = (mb["returns"] - v).pow(2).mean()
    sym_loss = 0
    if cfg.sym_loss:                                                            # ppo.py:197-202
        obs_perm_mat = perm_matrix(stacked_obs_permutation(cfg.obs_permutation, cfg.frame_stack))
        act_perm_mat = perm_matrix(cfg.act_permutation)
        mirror_obs = torch.matmul(mb["obs"], obs_perm_mat)
        mirror_act = policy_mean(p, mirror_obs)
        m_mirror_act = torch.matmul(mirror_act, act_perm_mat)
        sym_loss = (mu - m_mirror_act).pow(2).mean()
    lin_vel_loss = torch.nn.functional.mse_loss(lv, mb["lin_vel_target"])
    loss = (surrogate_loss + cfg.value_loss_coef * value_loss - cfg.entropy_coef * ent.mean() + cfg.sym_coef * sym_loss
            + cfg.base_lin_vel_coef * lin_vel_loss)
    # mean |term| of the surrogate: the scale its forward error is measured on (the loss itself is a near-cancelling mean
    # of signed terms, advantages being normalised to mean 0) -- used by the parity tests only
    surrogate_abs = torch.max(surrogate, surrogate_clipped).detach().abs().mean()
    return dict(loss=loss, surrogate=surrogate_loss, value=value_loss, lin_vel=lin_vel_loss, sym=sym_loss,
                entropy=ent.mean(), kl=kl_mean, mu=mu, v=v, lv=lv, surrogate_abs=surrogate_abs)


def clip_grad_norm(grads: List[torch.Tensor], max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (ppo.py:213): 2-norm of per-tensor 2-norms,
    coef = min(1, max_norm/(total+1e-6)), in place."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return float(total)


@dataclass
class AdamState:
    step: int = 0
    exp_avg: Dict[str, torch.Tensor] = field(default_factory=dict)
    exp_avg_sq: Dict[str, torch.Tensor] = field(default_factory=dict)


def adam_step(p: Params, grads: Dict[str, torch.Tensor], st: AdamState, lr: float, cfg: PPOConfig) -> None:
    """torch.optim.Adam defaults (no weight decay, no amsgrad), single-tensor formulation."""
    st.step += 1
    bc1 = 1.0 - cfg.beta1 ** st.step
    bc2 = 1.0 - cfg.beta2 ** st.step
    step_size = lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for k, g in grads.items():
        if k not in st.exp_avg:
            st.exp_avg[k] = torch.zeros_like(p[k])
            st.exp_avg_sq[k] = torch.zeros_like(p[k])
        m, v = st.exp_avg[k], st.exp_avg_sq[k]
        m.lerp_(g, 1.0 - cfg.beta1)
        v.mul_(cfg.beta2).addcmul_(g, g, value=1.0 - cfg.beta2)
        denom = (v.sqrt() / bc2_sqrt).add_(cfg.adam_eps)
        p[k].addcdiv_(m, denom, value=-step_size)


def ppo_update(p: Params, st: AdamState, ro: Rollout, perm: torch.Tensor, cfg: PPOConfig, lr: float,
               record: Optional[list] = None):
    """One `PPO.update()` (ppo.py:144-226).  Mutates p / st in place.
    Returns (mean_value_loss, mean_surrogate_loss, mean_lin_vel_loss, lr)."""
    keys = state_dict_order(p)
    sums = dict(value=0.0, surrogate=0.0, lin_vel=0.0)
    n = 0
    for idx in minibatch_slices(perm, cfg.num_mini_batches, cfg.num_learning_epochs):
        mb = gather_minibatch(ro, idx)
        leaves = {k: p[k].detach().requires_grad_(True) for k in keys}
        out = minibatch_losses(leaves, mb, cfg)
        if cfg.desired_kl is not None and cfg.schedule == "adaptive":
            lr = adaptive_lr(lr, float(out["kl"]), cfg.desired_kl)
        gl = torch.autograd.grad(out["loss"], [leaves[k] for k in keys], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(keys, gl)}
        gnorm = clip_grad_norm(list(grads.values()), cfg.max_grad_norm)
        adam_step(p, grads, st, lr, cfg)
        for k in sums:
            sums[k] += float(out[k].detach())
        n += 1
        if record is not None:
            record.append(dict(value=float(out["value"]), surrogate=float(out["surrogate"]),
                               lin_vel=float(out["lin_vel"]), entropy=float(out["entropy"]),
                               kl=float(out["kl"]), lr=lr, grad_norm=gnorm, sym=float(out["sym"]),
                               surrogate_abs=float(out["surrogate_abs"])))
    return sums["value"] / n, sums["surrogate"] / n, sums["lin_vel"] / n, lr


# ----------------------------------------------------------------------------- synthetic driver
def synthetic_rollout(p: Params, N: int, T: int, num_obs: int, num_priv: int, num_actions: int,
                      cfg: PPOConfig, gen: torch.Generator, record_inputs: bool = False):
    """BASELINE.json config 1/2 input distributions: obs, critic_obs ~ N(0,1),
    rewards ~ U(0,1), dones ~ Bernoulli(.02), time_outs ~ Bernoulli(.01), eps ~ N(0,1).
    Follows ppo.py:116-142 step by step."""
    obs = torch.zeros(T, N, num_obs)
    pobs = torch.zeros(T, N, num_priv)
    actions = torch.zeros(T, N, num_actions)
    rewards = torch.zeros(T, N, 1)
    dones = torch.zeros(T, N, 1, dtype=torch.uint8)
    values = torch.zeros(T, N, 1)
    logp = torch.zeros(T, N, 1)
    mu = torch.zeros(T, N, num_actions)
    sigma = torch.zeros(T, N, num_actions)
    inputs = []
    with torch.no_grad():
        for t in range(T):
            o = torch.randn(N, num_obs, generator=gen)
            c = torch.randn(N, num_priv, generator=gen)
            eps = torch.randn(N, num_actions, generator=gen)
            rew = torch.rand(N, generator=gen)
            dn = torch.rand(N, generator=gen) < 0.02
            to = torch.rand(N, generator=gen) < 0.01
            m = policy_mean(p, o)
            a = sample_actions(m, p["std"], eps)
            v = value(p, c)
            obs[t], pobs[t], actions[t], values[t] = o, c, a, v
            logp[t] = log_prob(m, p["std"], a).view(-1, 1)
            mu[t], sigma[t] = m, m * 0.0 + p["std"]
            rewards[t] = bootstrap_rewards(rew, v, to, cfg.gamma).view(-1, 1)
            dones[t] = dn.view(-1, 1)
            if record_inputs:
                inputs.append(dict(eps=eps, rew=rew, dones=dn, time_outs=to))
        last_c = torch.randn(N, num_priv, generator=gen)
        last_v = value(p, last_c)
    ro = Rollout(obs, pobs, actions, rewards, dones, values, logp, mu, sigma)
    ro.returns, ro.advantages = gae_returns(rewards, values, dones, last_v, cfg.gamma, cfg.lam)
    return ro, last_c, inputs
