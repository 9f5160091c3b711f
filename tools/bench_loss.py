"""Stand-alone timing of hg_ppo_head_loss at the update's minibatch shape (61440 rows, H=128, A=12); CUDA events around
the C call, so ~15 us of host launch latency are inside the figure (the device timeline below and ncu are not affected).
HG_TRACE_LOSS=1 with a trace build (ALL=1 tools/build_variant.sh trace -DHG_TRACE; HG_LIB_PATH=variants/libhgb200_trace.so)
prints the kernel's device timeline: when each block passed its phase boundaries, per tile round."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ctypes as C
import torch
from humanoid_gym_with_comments_b200._lib import HgHeadNet, HgLossArgs, check, current_stream, load, ptr
DEV = "cuda:0"
rows, A, H = int(os.environ.get("HG_ROWS", 61440)), 12, 128
g = torch.Generator(device=DEV).manual_seed(0)
TN = rows * 4
rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
h = [torch.nn.functional.elu(rn(rows, H)).to(torch.bfloat16) for _ in range(3)]
n_out = [A, 1, 3]
W = [0.1 * rn(n, H) for n in n_out]; b = [0.1 * rn(n) for n in n_out]
std = 0.7 + 0.5 * torch.rand(A, device=DEV, generator=g)
idx = torch.randperm(TN, device=DEV, generator=g)[:rows]
store = dict(actions=rn(TN, A), old_mu=0.3 * rn(TN, A), old_sigma=0.8 + 0.4 * torch.rand(TN, A, device=DEV, generator=g),
             old_logp=-12 + rn(TN), adv=rn(TN), ret=rn(TN), old_v=rn(TN), cobs=rn(TN, 219))
gW = [torch.zeros_like(w) for w in W]; gb = [torch.zeros_like(x) for x in b]
gbh = [torch.zeros(H, device=DEV) for _ in range(3)]
dZ = [torch.zeros(rows, H, dtype=torch.bfloat16, device=DEV) for _ in range(3)]
gstd = torch.zeros(A, device=DEV); stats = torch.zeros(8, device=DEV)
nets = [HgHeadNet(h=ptr(h[i]), ld_h=H, H=H, n_out=n_out[i], W=ptr(W[i]), b=ptr(b[i]), gW=ptr(gW[i]), gb=ptr(gb[i]),
                  dZ=ptr(dZ[i]), ld_dz=H, gb_hidden=ptr(gbh[i])) for i in range(3)]
args = HgLossArgs(actor=nets[0], critic=nets[1], linvel=nets[2], rows=rows, A=A, idx=ptr(idx),
                  actions=ptr(store["actions"]), old_mu=ptr(store["old_mu"]), old_sigma=ptr(store["old_sigma"]),
                  old_log_prob=ptr(store["old_logp"]), advantages=ptr(store["adv"]), returns=ptr(store["ret"]),
                  old_values=ptr(store["old_v"]), critic_obs=ptr(store["cobs"]), ld_cobs=219, lv_col=53, std=ptr(std),
                  g_std=ptr(gstd), clip_param=0.2, value_loss_coef=1.0, entropy_coef=0.001, lin_vel_coef=1.0,
                  use_clipped_value_loss=1, inv_world=1.0, stats=ptr(stats))
lib = load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
ts = []
for i in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); check(lib.hg_ppo_head_loss(C.byref(args), current_stream())); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("loss head", "SIMT" if os.environ.get("HG_LOSS_SIMT") else "mma", f"rows {rows}: median {sorted(ts[2:])[3]:.1f} us  all {[round(t, 1) for t in ts]}")

if hasattr(lib, "hg_trace_set") and os.environ.get("HG_TRACE_LOSS"):
    # device timeline of one launch (trace build: ALL=1 tools/build_variant.sh trace -DHG_TRACE): per tile round, when each
    # block reached the phase boundaries (ns after the launch's first record; min / median / max over blocks)
    TB = 512
    table = torch.zeros(4 * 64 * 4 * TB, dtype=torch.int64, device=DEV)
    lib.hg_trace_set.restype = C.c_int
    lib.hg_trace_set.argtypes = [C.c_void_p]
    flush.zero_(); torch.cuda.synchronize()
    assert lib.hg_trace_set(table.data_ptr()) == 0
    check(lib.hg_ppo_head_loss(C.byref(args), current_stream())); torch.cuda.synchronize()
    assert lib.hg_trace_set(None) == 0
    hh = table.cpu().view(4, 64, 4, TB)[0]
    t0 = int(hh[hh > 0].min())
    names = {0: "kernel entry", 8: "next tile issued", 1: "prologue done", 2: "tile start", 3: "activations landed", 4: "forward done", 5: "loss done", 6: "backward done", 7: "kernel end"}
    for rnd in range(4):
        for pt in (2, 8, 3, 4, 5, 6):
            v = hh[pt, rnd]; v = sorted((v[v > 0] - t0).tolist())
            if v: print(f"round {rnd} {names[pt]:20s} n={len(v):4d} min {v[0]:7d} med {v[len(v)//2]:7d} max {v[-1]:7d}")
    for pt in (0, 1, 7):
        v = hh[pt, 0]; v = sorted((v[v > 0] - t0).tolist())
        print(f"{names[pt]:28s} n={len(v):4d} min {v[0]:7d} med {v[len(v)//2]:7d} max {v[-1]:7d}")
    # per-block phase durations (median over blocks and rounds)
    import statistics
    for a, b_, nm in ((2, 8, "issue next"), (8, 3, "load wait"), (3, 4, "forward"), (4, 5, "loss"), (5, 6, "backward")):
        d = []
        for rnd in range(4):
            x, y = hh[a, rnd], hh[b_, rnd]
            m = (x > 0) & (y > 0)
            d += (y[m] - x[m]).tolist()
        print(f"phase {nm:10s}: median {statistics.median(d):.0f} ns  mean {statistics.mean(d):.0f}  max {max(d)}")
