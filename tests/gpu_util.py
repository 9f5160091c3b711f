"""Helpers shared by the -m gpu tests: thin wrappers that call the C ABI through ctypes."""
import ctypes as C

import torch

from humanoid_gym_with_comments_b200 import _lib
from humanoid_gym_with_comments_b200._lib import HgGemmProblem, check, current_stream, ptr

DEV = "cuda:0"


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def make_problem(A, B, out, M, N, K, a_mn=0, b_mn=0, k_splits=1, bias=None, act=None, colsum=None):
    return HgGemmProblem(A=ptr(A), B=ptr(B), lda=A.stride(0), ldb=B.stride(0), a_mn_major=a_mn, b_mn_major=b_mn,
                         M=M, N=N, K=K, k_splits=k_splits, out=ptr(out), ldo=out.stride(0), bias=ptr(bias),
                         act=ptr(act), ldact=act.stride(0) if act is not None else 0, colsum=ptr(colsum))


def gemm(problems, epi, block_n, cg=1):
    arr = (HgGemmProblem * len(problems))(*problems)
    check(_lib.load().hg_gemm_grouped(arr, len(problems), epi, block_n, cg, current_stream()))
    torch.cuda.synchronize()


def padded_bf16(rows, cols, ld=None, scale=1.0, gen=None):
    """bf16 [rows, ld] with random data in the first `cols` columns and NaN poison in the padding
    (TMA must never read it: the tensor map's extent is `cols`)."""
    ld = ld or (cols + 7) // 8 * 8
    t = torch.full((rows, ld), float("nan"), dtype=torch.bfloat16, device=DEV)
    t[:, :cols] = (scale * torch.randn(rows, cols, device=DEV, generator=gen)).to(torch.bfloat16)
    return t


def report(name, got, want, atol, rtol):
    err = (got - want).abs()
    tol = atol + rtol * want.abs()
    bad = ~(err <= tol)          # NaN counts as a mismatch
    nbad = int(bad.sum())
    msg = f"{name}: max_abs_err={float(err.max()):.3e} max_ref={float(want.abs().max()):.3e} bad={nbad}/{bad.numel()}"
    if nbad:
        idx = bad.nonzero()[:8].tolist()
        msg += " first_bad=" + str([(i, float(got[tuple(i)]), float(want[tuple(i)])) for i in idx])
    print(msg)
    return nbad == 0, msg
