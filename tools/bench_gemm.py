"""Micro-benchmark of the grouped tcgen05 GEMM on the hot path's shapes (CUDA-event timed, L2 flushed between reps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import DEV, gemm, make_problem
from humanoid_gym_with_comments_b200 import _lib

M = int(os.environ.get("HG_M", 61440))
CG = int(os.environ.get("HG_CG", 1))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


def bf(rows, cols, scale=1.0):
    return (scale * torch.randn(rows, cols, device=DEV)).to(torch.bfloat16)


def timeit(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


def fwd(specs, bn, m=M):
    probs, keep, fl = [], [], 0
    for N, K in specs:
        al = int(os.environ.get('HG_LD_ALIGN', 8))
        ld = (K + al - 1) // al * al
        A, W = bf(m, ld), bf(N, ld, 0.05)
        out = torch.empty(m, N, dtype=torch.bfloat16, device=DEV)
        bias = torch.zeros(N, device=DEV)
        probs.append(make_problem(A, W, out, m, N, K, bias=bias)); keep.append((A, W, out, bias)); fl += 2.0 * m * N * K
    ms = timeit(lambda: gemm_nosync(probs, _lib.EPI_BIAS_ELU_BF16, bn))
    return ms, fl


def dx(specs, bn, m=M):
    probs, keep, fl = [], [], 0
    for NL, KL in specs:      # dZ [m, NL] @ W [NL, KL]
        dZ, W = bf(m, NL, 1e-3), bf(NL, (KL + 7) // 8 * 8, 0.05)
        H = bf(m, KL); out = torch.empty(m, KL, dtype=torch.bfloat16, device=DEV); cs = torch.zeros(KL, device=DEV)
        probs.append(make_problem(dZ, W, out, m, KL, NL, 0, 1, act=H, colsum=cs)); keep.append((dZ, W, H, out, cs)); fl += 2.0 * m * NL * KL
    ms = timeit(lambda: gemm_nosync(probs, _lib.EPI_DELU_BF16, bn))
    return ms, fl


def dw(specs, bn, splits, m=M):
    probs, keep, fl = [], [], 0
    for NL, KL in specs:      # dZ^T [NL, m] @ H [m, KL]
        dZ, H = bf(m, NL, 1e-3), bf(m, (KL + 7) // 8 * 8)
        out = torch.zeros(NL, KL, device=DEV)
        probs.append(make_problem(dZ, H, out, NL, KL, m, 1, 1, k_splits=splits)); keep.append((dZ, H, out)); fl += 2.0 * m * NL * KL
    ms = timeit(lambda: gemm_nosync(probs, _lib.EPI_ATOMIC_F32, bn))
    return ms, fl


def gemm_nosync(problems, epi, bn):
    from humanoid_gym_with_comments_b200._lib import HgGemmProblem, check, current_stream
    arr = (HgGemmProblem * len(problems))(*problems)
    check(_lib.load().hg_gemm_grouped(arr, len(problems), epi, bn, CG, current_stream()))


def show(name, ms, fl):
    print(f"cg{CG} {name:55s} {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    for bn in (128, 256):
        show(f"fwd L1 a+c+l bn{bn}", *fwd([(512, 705), (768, 219), (128, 705)], bn))
        show(f"fwd L1 actor only bn{bn}", *fwd([(512, 705)], bn))
        show(f"fwd L2 a+c+l bn{bn}", *fwd([(256, 512), (256, 768), (128, 128)], bn))
        show(f"fwd L3 a+c bn{bn}", *fwd([(128, 256), (128, 256)], bn))
        show(f"dx L3 group bn{bn}", *dx([(128, 256), (128, 256), (128, 128)], bn))
        show(f"dx L2 group bn{bn}", *dx([(256, 512), (256, 768)], bn))
        for sp in (4, 8, 16, 29):
            show(f"dw L1 group bn{bn} splits{sp}", *dw([(512, 705), (768, 219)], bn, sp))
        show(f"dw L3 group bn{bn} splits29", *dw([(128, 256), (128, 256), (128, 128)], bn, 29))
        show(f"dw L2 group bn{bn} splits12", *dw([(256, 512), (256, 768), (128, 705)], bn, 12))
        show(f"big square 8192^3 bn{bn}", *fwd([(8192, 8192)], bn, m=8192))
    for m in (4096,):
        for bn in (128, 256):
            show(f"rollout fwd L1 a+c M{m} bn{bn}", *fwd([(512, 705), (768, 219)], bn, m))
            show(f"rollout fwd L2 a+c M{m} bn{bn}", *fwd([(256, 512), (256, 768)], bn, m))
            show(f"rollout fwd L3 a+c M{m} bn{bn}", *fwd([(128, 256), (128, 256)], bn, m))
