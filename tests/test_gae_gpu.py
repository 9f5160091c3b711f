"""GAE scan kernel vs the oracle restatement of rollout_storage.py:122-143 (-m gpu)."""
import numpy as np
import pytest
import torch

from gpu_util import DEV
from humanoid_gym_with_comments_b200 import _lib
from humanoid_gym_with_comments_b200._lib import check, current_stream, ptr
from oracle import ppo_oracle as po

pytestmark = pytest.mark.gpu


def run_gae(rew, val, dones, last, gamma, lam):
    T, N = rew.shape[:2]
    d = lambda x: x.to(DEV).contiguous()
    r, v, dn, lv = d(rew), d(val), d(dones), d(last)
    ret = torch.full((T, N, 1), float("nan"), device=DEV)
    adv = torch.full((T, N, 1), float("nan"), device=DEV)
    scratch = torch.zeros(4, dtype=torch.float64, device=DEV)
    check(_lib.load().hg_gae_returns(ptr(r), ptr(v), ptr(dn), ptr(lv), gamma, lam, T, N, ptr(ret), ptr(adv), ptr(scratch),
                                     current_stream()))
    torch.cuda.synchronize()
    return ret.cpu(), adv.cpu()


@pytest.mark.parametrize("T,N", [(1, 1), (3, 5), (24, 64), (60, 4096), (17, 1000), (60, 129)])
def test_gae_matches_oracle(T, N):
    g = torch.Generator().manual_seed(T * 1000 + N)
    rew = torch.rand(T, N, 1, generator=g)
    val = torch.randn(T, N, 1, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < 0.05).to(torch.uint8)
    last = torch.randn(N, 1, generator=g)
    want_ret, want_adv = po.gae_returns(rew, val, dones, last, 0.994, 0.9)
    ret, adv = run_gae(rew, val, dones, last, 0.994, 0.9)
    assert torch.equal(ret, want_ret), f"returns differ: max {float((ret - want_ret).abs().max())}"    # bit-exact
    if T * N > 1:
        # advantages: same fp32 values up to the rounding of the global mean / std reduction; tolerance 1e-5 relative
        assert torch.allclose(adv, want_adv, rtol=1e-5, atol=2e-6), float((adv - want_adv).abs().max())


def test_gae_golden_fixture(golden_dir):
    z = np.load(f"{golden_dir}/learner_small.npz")
    rew, val = torch.tensor(z["u0.rewards"]), torch.tensor(z["u0.values"])
    dones = torch.tensor(z["u0.dones"])
    last = torch.tensor(z["u0.last_values"])
    ret, adv = run_gae(rew, val, dones, last, 0.994, 0.9)
    assert torch.equal(ret, torch.tensor(z["u0.returns"]))
    assert torch.allclose(adv, torch.tensor(z["u0.advantages"]), rtol=1e-5, atol=2e-6)


def test_gae_all_done_and_no_done():
    T, N = 8, 33
    rew, val, last = torch.ones(T, N, 1), torch.zeros(T, N, 1), torch.ones(N, 1)
    for dones in (torch.ones(T, N, 1, dtype=torch.uint8), torch.zeros(T, N, 1, dtype=torch.uint8)):
        want_ret, _ = po.gae_returns(rew, val, dones, last, 0.9, 0.8)
        ret, _ = run_gae(rew, val, dones, last, 0.9, 0.8)
        assert torch.equal(ret, want_ret)
