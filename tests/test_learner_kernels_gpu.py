"""Unit parity of the learner kernels against a plain torch fp32 reference of the same op (-m gpu):
rollout head, fused head+loss+head-backward, clip+Adam(+adaptive lr), ingest, gather."""
import ctypes as C
import math

import pytest
import torch

from gpu_util import DEV
from humanoid_gym_with_comments_b200 import _lib
from humanoid_gym_with_comments_b200._lib import (HgAdamArgs, HgBf16Shadow, HgHeadNet, HgLossArgs, check, current_stream,
                                                  ptr)

pytestmark = pytest.mark.gpu
lib = _lib.load


def test_ingest_and_gather():
    g = torch.Generator(device=DEV).manual_seed(0)
    rows, cols, ld = 300, 705, 768
    src = torch.randn(rows, cols, device=DEV, generator=g)
    dst = torch.zeros(rows, cols, device=DEV)
    d16 = torch.full((rows, ld), float("nan"), dtype=torch.bfloat16, device=DEV)
    check(lib().hg_ingest_rows(ptr(src), cols, rows, cols, ptr(dst), cols, ptr(d16), ld, current_stream()))
    assert torch.equal(dst, src)
    assert torch.equal(d16[:, :cols], src.to(torch.bfloat16))
    assert torch.all(d16[:, cols:] == 0)
    idx = torch.randperm(rows, device=DEV, generator=g)[:128]
    out = torch.zeros(128, ld, dtype=torch.bfloat16, device=DEV)
    check(lib().hg_gather_rows_bf16(ptr(d16), ld, ptr(idx), 128, ptr(out), current_stream()))
    assert torch.equal(out, d16[idx])          # bit-exact row indexing


@pytest.mark.parametrize("n,A,H", [(4096, 12, 128), (100, 18, 256), (33, 12, 64)])
def test_rollout_head(n, A, H):
    g = torch.Generator(device=DEV).manual_seed(n)
    ha = torch.randn(n, H, device=DEV, generator=g).to(torch.bfloat16)
    hc = torch.randn(n, H, device=DEV, generator=g).to(torch.bfloat16)
    Wa, ba = 0.1 * torch.randn(A, H, device=DEV, generator=g), 0.1 * torch.randn(A, device=DEV, generator=g)
    Wc, bc = 0.1 * torch.randn(1, H, device=DEV, generator=g), 0.1 * torch.randn(1, device=DEV, generator=g)
    std = 0.5 + torch.rand(A, device=DEV, generator=g)
    eps = torch.randn(n, A, device=DEV, generator=g)
    act, mu, sig = (torch.zeros(n, A, device=DEV) for _ in range(3))
    lp, val = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    check(lib().hg_policy_head_rollout(ptr(ha), H, H, ptr(Wa), ptr(ba), ptr(std), ptr(hc), H, H, ptr(Wc), ptr(bc), ptr(eps),
                                       0, 0, n, A, ptr(act), ptr(mu), ptr(sig), ptr(lp), ptr(val), current_stream()))
    mu_ref = ha.float() @ Wa.t() + ba
    act_ref = mu_ref + std * eps
    dist = torch.distributions.Normal(mu_ref, mu_ref * 0 + std)
    assert torch.allclose(mu, mu_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(act, act_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(sig, (mu_ref * 0 + std))
    assert torch.allclose(lp, dist.log_prob(act_ref).sum(-1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(val, (hc.float() @ Wc.t() + bc).squeeze(1), rtol=1e-5, atol=1e-5)
    # Philox path: unit-variance, zero-mean draws, different per row
    check(lib().hg_policy_head_rollout(ptr(ha), H, H, ptr(Wa), ptr(ba), ptr(std), ptr(hc), H, H, ptr(Wc), ptr(bc), 0,
                                       123, 7, n, A, ptr(act), ptr(mu), ptr(sig), ptr(lp), ptr(val), current_stream()))
    z = (act - mu) / std
    if n >= 1000:
        assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.02


def _loss_reference(h, W, b, std, batch, cfg):
    """torch fp32 autograd reference of the head+loss op, from the same bf16 activations."""
    ha, hc, hl = (x.float().requires_grad_(True) for x in h)
    Wa, Wc, Wl = (x.clone().requires_grad_(True) for x in W)
    ba, bc, bl = (x.clone().requires_grad_(True) for x in b)
    std = std.clone().requires_grad_(True)
    mu = ha @ Wa.t() + ba
    v = hc @ Wc.t() + bc
    lv = hl @ Wl.t() + bl
    sigma = mu * 0 + std
    dist = torch.distributions.Normal(mu, sigma)
    logp = dist.log_prob(batch["actions"]).sum(-1)
    ent = dist.entropy().sum(-1)
    kl = torch.sum(torch.log(sigma / batch["old_sigma"] + 1e-5) + (batch["old_sigma"] ** 2 + (batch["old_mu"] - mu) ** 2)
                   / (2 * sigma ** 2) - 0.5, -1)
    ratio = torch.exp(logp - batch["old_logp"])
    adv = batch["adv"]
    surr = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - cfg["clip"], 1 + cfg["clip"])).mean()
    vc = batch["old_v"] + (v.squeeze(1) - batch["old_v"]).clamp(-cfg["clip"], cfg["clip"])
    vl = torch.max((v.squeeze(1) - batch["ret"]) ** 2, (vc - batch["ret"]) ** 2).mean()
    ll = torch.nn.functional.mse_loss(lv, batch["lv_tgt"])
    loss = surr + cfg["vcoef"] * vl - cfg["ecoef"] * ent.mean() + cfg["lcoef"] * ll
    loss.backward()
    elu_d = lambda hh: torch.where(hh > 0, torch.ones_like(hh), hh + 1)
    return dict(stats=[vl.item(), surr.item(), ll.item(), kl.mean().item(), ent.mean().item()],
                gW=[Wa.grad, Wc.grad, Wl.grad], gb=[ba.grad, bc.grad, bl.grad], gstd=std.grad,
                dZ=[ha.grad * elu_d(ha.detach()), hc.grad * elu_d(hc.detach()), hl.grad * elu_d(hl.detach())])


@pytest.mark.parametrize("rows,A,H", [(384, 12, 128), (1000, 18, 128), (61440, 12, 128), (200, 12, 64)])
def test_head_loss_forward_backward(rows, A, H):
    g = torch.Generator(device=DEV).manual_seed(rows + A)
    TN = rows * 2
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    h = [torch.nn.functional.elu(rn(rows, H)).to(torch.bfloat16) for _ in range(3)]
    n_out = [A, 1, 3]
    W = [0.1 * rn(n, H) for n in n_out]
    b = [0.1 * rn(n) for n in n_out]
    std = 0.7 + 0.5 * torch.rand(A, device=DEV, generator=g)
    idx = torch.randperm(TN, device=DEV, generator=g)[:rows]
    store = dict(actions=rn(TN, A), old_mu=0.3 * rn(TN, A), old_sigma=0.8 + 0.4 * torch.rand(TN, A, device=DEV, generator=g),
                 old_logp=-12 + rn(TN), adv=rn(TN), ret=rn(TN), old_v=rn(TN), cobs=rn(TN, 219))
    cfg = dict(clip=0.2, vcoef=1.0, ecoef=0.001, lcoef=1.0)
    # make the ratio hit both clip branches
    with torch.no_grad():
        mu0 = h[0].float() @ W[0].t() + b[0]
        lp0 = torch.distributions.Normal(mu0, mu0 * 0 + std).log_prob(store["actions"][idx]).sum(-1)
        store["old_logp"][idx] = lp0 + 0.3 * rn(rows)
    batch = dict(actions=store["actions"][idx], old_mu=store["old_mu"][idx], old_sigma=store["old_sigma"][idx],
                 old_logp=store["old_logp"][idx], adv=store["adv"][idx], ret=store["ret"][idx], old_v=store["old_v"][idx],
                 lv_tgt=store["cobs"][idx][:, 53:56])
    ref = _loss_reference(h, W, b, std, batch, cfg)

    gW = [torch.zeros_like(w) for w in W]
    gb = [torch.zeros_like(x) for x in b]
    gbh = [torch.zeros(H, device=DEV) for _ in range(3)]
    dZ = [torch.zeros(rows, H, dtype=torch.bfloat16, device=DEV) for _ in range(3)]
    gstd = torch.zeros(A, device=DEV)
    stats = torch.zeros(8, device=DEV)
    nets = [HgHeadNet(h=ptr(h[i]), ld_h=H, H=H, n_out=n_out[i], W=ptr(W[i]), b=ptr(b[i]), gW=ptr(gW[i]), gb=ptr(gb[i]),
                      dZ=ptr(dZ[i]), ld_dz=H, gb_hidden=ptr(gbh[i])) for i in range(3)]
    args = HgLossArgs(actor=nets[0], critic=nets[1], linvel=nets[2], rows=rows, A=A, idx=ptr(idx),
                      actions=ptr(store["actions"]), old_mu=ptr(store["old_mu"]), old_sigma=ptr(store["old_sigma"]),
                      old_log_prob=ptr(store["old_logp"]), advantages=ptr(store["adv"]), returns=ptr(store["ret"]),
                      old_values=ptr(store["old_v"]), critic_obs=ptr(store["cobs"]), ld_cobs=219, lv_col=53, std=ptr(std),
                      g_std=ptr(gstd), clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.001, lin_vel_coef=1.0,
                      use_clipped_value_loss=1, inv_world=1.0, stats=ptr(stats))
    check(lib().hg_ppo_head_loss(C.byref(args), current_stream()))
    torch.cuda.synchronize()
    got = [stats[0].item() / rows, stats[1].item() / rows, stats[2].item() / (3 * rows), stats[3].item() / rows,
           stats[4].item() / rows]
    for name, a, r in zip(("value", "surrogate", "lin_vel", "kl", "entropy"), got, ref["stats"]):
        assert abs(a - r) <= 2e-5 * max(1.0, abs(r)), (name, a, r)
    for i in range(3):
        assert torch.allclose(gW[i], ref["gW"][i], rtol=2e-4, atol=1e-6), (i, float((gW[i] - ref["gW"][i]).abs().max()))
        assert torch.allclose(gb[i], ref["gb"][i], rtol=2e-4, atol=1e-6), i
        # dZ is stored in bf16: 2^-8 relative; the tensor-core path forms dOut W from 16-bit split operands, so an
        # element that is a near-cancelling sum carries ~2^-16 of its TERMS (hence the scale-aware atol)
        atol_dz = 2e-5 * float(ref["dZ"][i].abs().max())
        assert torch.allclose(dZ[i].float(), ref["dZ"][i], rtol=1e-2, atol=atol_dz), (i, float((dZ[i].float() - ref["dZ"][i]).abs().max()))
        assert torch.allclose(gbh[i], ref["dZ"][i].sum(0), rtol=1e-3, atol=1e-6), i
    assert torch.allclose(gstd, ref["gstd"], rtol=2e-4, atol=1e-6)


def test_clip_adam_matches_torch():
    """3 consecutive steps vs torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the same gradients, with the
    adaptive-KL learning-rate rule of ppo.py:169-172 driven through the statistics slots."""
    g = torch.Generator(device=DEV).manual_seed(0)
    shapes = [(12,), (512, 705), (512,), (12, 128), (1, 128)]
    count = sum(((math.prod(s) + 3) // 4) * 4 for s in shapes)
    P = torch.zeros(count + 8, device=DEV)
    G, M1, M2 = torch.zeros_like(P), torch.zeros_like(P), torch.zeros_like(P)
    views, offs, o = [], [], 0
    for s in shapes:
        n = math.prod(s)
        views.append(P[o:o + n].view(s))
        offs.append(o)
        o += ((n + 3) // 4) * 4
    for v in views:
        v.copy_(0.1 * torch.randn(v.shape, device=DEV, generator=g))
    ref_params = [torch.nn.Parameter(v.clone()) for v in views]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    sh = torch.zeros(512, 768, dtype=torch.bfloat16, device=DEV)
    shadows = (HgBf16Shadow * 1)(HgBf16Shadow(offs[1], 512, 705, sh.data_ptr(), 768))
    opt_state = torch.zeros(8, dtype=torch.float64, device=DEV)
    opt_state[0] = 1e-3
    log = torch.zeros(3, 8, dtype=torch.float64, device=DEV)
    lr = 1e-3
    rows = 1000
    for step, kl_mean in enumerate([0.001, 0.05, 0.01]):
        grads = [(5.0 if step == 0 else 0.01) * torch.randn(v.shape, device=DEV, generator=g) for v in views]
        for s_, o_, gr in zip(shapes, offs, grads):
            G[o_:o_ + gr.numel()].copy_(gr.reshape(-1))
        G[count + 3] = kl_mean * rows
        args = HgAdamArgs(params=ptr(P), grads=ptr(G), exp_avg=ptr(M1), exp_avg_sq=ptr(M2), count=count,
                          opt_state=ptr(opt_state), loss_log=ptr(log), log_slot=step, rows_global=rows, beta1=0.9,
                          beta2=0.999, eps=1e-8, max_grad_norm=1.0, adaptive=1, desired_kl=0.01, lr_min=1e-5, lr_max=1e-2,
                          shadows_host=shadows, num_shadows=1, zero_grads=1)
        check(lib().hg_clip_adam_step(C.byref(args), current_stream()))
        if kl_mean > 0.02:
            lr = max(1e-5, lr / 1.5)
        elif kl_mean < 0.005 and kl_mean > 0:
            lr = min(1e-2, lr * 1.5)
        for pg in opt.param_groups:
            pg["lr"] = lr
        for p_, gr in zip(ref_params, grads):
            p_.grad = gr.clone()
        total = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        opt.step()
        torch.cuda.synchronize()
        assert abs(log[step, 5].item() - lr) < 1e-12
        assert abs(log[step, 6].item() - total.item()) <= 1e-5 * total.item()
        for v, p_ in zip(views, ref_params):
            assert torch.allclose(v, p_.data, rtol=1e-5, atol=1e-7), float((v - p_.data).abs().max())
        assert torch.all(G == 0)
        assert torch.equal(sh[:, :705], views[1].to(torch.bfloat16)) and torch.all(sh[:, 705:] == 0)
    assert opt_state[1].item() == 3
