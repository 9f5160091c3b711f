"""Device timeline of ONE update-sized grouped-GEMM launch from a trace build (see tools/trace_rollout.py):
    HG_LIB_PATH=$PWD/variants/libhgb200_trace.so python tools/trace_gemm.py [dx1|dx2|fwd1|fwd3]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench_gemm as bg
import bench_gemm_dx as bd
from gpu_util import make_problem
from humanoid_gym_with_comments_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "dx1"
lib = _lib.load()
TB = 512
table = torch.zeros(4 * 64 * 4 * TB, dtype=torch.int64, device=bg.DEV)
lib.hg_trace_set.restype = C.c_int
lib.hg_trace_set.argtypes = [C.c_void_p]
m = bg.M
if which in ("dx1", "dx2"):
    specs = [(256, 512), (256, 768), (128, 128)] if which == "dx1" else [(128, 256), (128, 256)]
    probs, keep = [], []
    for NL, KL in specs:
        dZ, W = bg.bf(m, NL, 1e-3), bg.bf(NL, KL, 0.05)
        H = bg.bf(m, KL); out = torch.empty(m, KL, dtype=torch.bfloat16, device=bg.DEV); cs = torch.zeros(KL, device=bg.DEV)
        probs.append(make_problem(dZ, W, out, m, KL, NL, 0, 1, act=H, colsum=cs)); keep.append((dZ, W, H, out, cs))
    run = lambda: bg.gemm_nosync(probs, _lib.EPI_DELU_BF16, 256)
elif which in ("dw3", "dw2"):
    # dW = dZ^T [NL, m] @ H [m, KL]: split-K over the rows, fp32 atomic epilogue (layer 3 / layer 2 shapes)
    specs = [(128, 256), (128, 256), (128, 128)] if which == "dw3" else [(256, 512), (256, 768), (128, 720)]
    bg.CG = int(os.environ.get("HG_CG", 2))
    tiles = sum(((NL + 128 * bg.CG - 1) // (128 * bg.CG)) * ((KL + 255) // 256) for NL, KL in specs)
    splits = max(1, min((m + 63) // 64, (148 // bg.CG) // max(tiles, 1)))
    probs, keep = [], []
    for NL, KL in specs:
        dZ, H = bg.bf(m, NL, 1e-3), bg.bf(m, (KL + 63) // 64 * 64)
        out = torch.zeros(NL, KL, device=bg.DEV)
        probs.append(make_problem(dZ, H, out, NL, KL, m, 1, 1, k_splits=splits)); keep.append((dZ, H, out))
    print("splits", splits, "tiles", tiles, "cg", bg.CG)
    run = lambda: bg.gemm_nosync(probs, _lib.EPI_ATOMIC_F32, 256)
else:
    specs = [(512, 720), (768, 240), (128, 720)] if which == "fwd1" else [(128, 256), (128, 256), (128, 128)]
    probs, keep = [], []
    for N, K in specs:
        ld = (K + 63) // 64 * 64
        A, W = bg.bf(m, ld), bg.bf(N, ld, 0.05)
        out = torch.empty(m, N, dtype=torch.bfloat16, device=bg.DEV); bias = torch.zeros(N, device=bg.DEV)
        probs.append(make_problem(A, W, out, m, N, K, bias=bias)); keep.append((A, W, out, bias))
    bg.CG = 2 if which == "fwd1" else bg.CG
    run = lambda: bg.gemm_nosync(probs, _lib.EPI_BIAS_ELU_BF16, 256 if which == "fwd1" else 128)
for _ in range(3):
    run()
torch.cuda.synchronize()
print(which, "untraced:", bg.timeit(run), "ms")
table.zero_(); torch.cuda.synchronize()
assert lib.hg_trace_set(table.data_ptr()) == 0
run(); torch.cuda.synchronize()
assert lib.hg_trace_set(None) == 0
h = table.cpu().view(4, 64, 4, TB)
t0 = int(h[h > 0].min())
NAMES = {0x200: "entry", 0x201: "set-up done", 0x202: "first TMA issued", 0x203: "first operands landed", 0x204: "tile MMAs committed",
         0x205: "epilogue sees accumulator", 0x208: "chunk in registers (tile*2+chunk)", 0x209: "chunk done (tile*2+chunk)",
         0x206: "epilogue tile done", 0x207: "end"}
rows = []
for tag, name in NAMES.items():
    k, pt = (tag >> 8) & 3, tag & 63
    for sub in range(4):
        v = h[k, pt, sub]
        v = sorted((v[v > 0] - t0).tolist())
        if v:
            rows.append((v[len(v) // 2], f"  {name:36s} [{sub}] n={len(v):4d}  min {v[0]:7d}  med {v[len(v) // 2]:7d}  max {v[-1]:7d}"))
for _, line in sorted(rows):
    print(line)
