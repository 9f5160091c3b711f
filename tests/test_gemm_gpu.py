"""tcgen05 grouped GEMM vs a plain torch fp32 matmul of the same bf16-rounded operands (-m gpu).
Covers both operand majors, both tile widths, ragged sizes (TMA out-of-bounds fill), split-K atomics,
the fused epilogues and grouped launches."""
import pytest
import torch

from gpu_util import DEV, bf16_round, gemm, make_problem, padded_bf16, report
from humanoid_gym_with_comments_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (128, 128, 128), (256, 256, 192), (384, 512, 705), (200, 96, 100), (4096, 768, 219),
          (1000, 130, 64)]


def ref_mm(A, B, a_mn, b_mn, M, N, K):
    a = A.float()[:K, :M].t() if a_mn else A.float()[:M, :K]
    b = B.float()[:K, :N].t() if b_mn else B.float()[:N, :K]
    return a @ b.t()


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_store_f32(M, N, K, a_mn, b_mn, block_n, cg):
    g = torch.Generator(device=DEV).manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = padded_bf16(K, M, gen=g) if a_mn else padded_bf16(M, K, gen=g)
    B = padded_bf16(K, N, gen=g) if b_mn else padded_bf16(N, K, gen=g)
    out = torch.full((M, N), float("nan"), device=DEV)
    gemm([make_problem(A, B, out, M, N, K, a_mn, b_mn)], _lib.EPI_STORE_F32, block_n, cg)
    ok, msg = report(f"store M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} bn{block_n} cg{cg}", out, ref_mm(A, B, a_mn, b_mn, M, N, K),
                     atol=1e-3 * K ** 0.5, rtol=1e-4)
    assert ok, msg


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
def test_bias_elu_forward(block_n, cg):
    M, N, K = 4096, 512, 705
    g = torch.Generator(device=DEV).manual_seed(1)
    A, W = padded_bf16(M, K, gen=g), padded_bf16(N, K, scale=0.04, gen=g)
    bias = 0.1 * torch.randn(N, device=DEV, generator=g)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    gemm([make_problem(A, W, out, M, N, K, bias=bias)], _lib.EPI_BIAS_ELU_BF16, block_n, cg)
    want = torch.nn.functional.elu(A.float()[:, :K] @ W.float()[:, :K].t() + bias)
    ok, msg = report("bias_elu", out.float(), want, atol=2e-3, rtol=1e-2)
    assert ok, msg


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
def test_delu_backward_with_colsum(block_n, cg):
    # dZ_prev = (dZ [M,N_l] @ W [N_l, K_l]) * elu'(H_prev),   colsum = bias gradient
    M, NL, KL = 2048, 256, 512
    g = torch.Generator(device=DEV).manual_seed(2)
    dZ = padded_bf16(M, NL, scale=1e-3, gen=g)
    W = padded_bf16(NL, KL, scale=0.05, gen=g)
    H = torch.nn.functional.elu(torch.randn(M, KL, device=DEV, generator=g)).to(torch.bfloat16)
    out = torch.zeros(M, KL, dtype=torch.bfloat16, device=DEV)
    colsum = torch.zeros(KL, device=DEV)
    gemm([make_problem(dZ, W, out, M, KL, NL, a_mn=0, b_mn=1, act=H, colsum=colsum)], _lib.EPI_DELU_BF16, block_n, cg)
    hf = H.float()
    want = (dZ.float() @ W.float()) * torch.where(hf > 0, torch.ones_like(hf), hf + 1)
    ok1, m1 = report("delu", out.float(), want, atol=1e-6, rtol=1e-2)
    # the bias gradient is the column sum of the STORED (bf16-rounded) dZ, accumulated in fp32 on the tensor core: exact
    # against the stored tensor up to summation order, and within bf16 rounding noise (2^-9 * |dZ| * sqrt(M)) of the fp32 sum
    ok2, m2 = report("colsum vs stored dZ", colsum, out.float().sum(0), atol=2e-5, rtol=1e-4)
    ok3, m3 = report("colsum vs fp32", colsum, want.sum(0), atol=5e-4, rtol=1e-3)
    assert ok1 and ok2 and ok3, " | ".join((m1, m2, m3))


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("splits", [1, 5, 16])
def test_weight_gradient_split_k(splits, block_n, cg):
    # dW [N_l, K_l] = dZ^T [N_l, M] @ H [M, K_l], both operands read MN-major from row-major [M, .] tensors
    M, NL, KL = 6144, 512, 705
    g = torch.Generator(device=DEV).manual_seed(3)
    dZ = padded_bf16(M, NL, scale=1e-2, gen=g)
    H = padded_bf16(M, KL, gen=g)
    out = torch.zeros(NL, KL, device=DEV)
    gemm([make_problem(dZ, H, out, NL, KL, M, a_mn=1, b_mn=1, k_splits=splits)], _lib.EPI_ATOMIC_F32, block_n, cg)
    want = dZ.float().t() @ H.float()[:, :KL]
    ok, msg = report(f"dW splits={splits}", out, want, atol=1e-4, rtol=1e-3)
    assert ok, msg


@pytest.mark.parametrize("cg", [1, 2])
def test_grouped_three_problems(cg):
    g = torch.Generator(device=DEV).manual_seed(4)
    M = 1024
    specs = [(512, 705), (768, 219), (128, 705)]
    probs, outs, wants = [], [], []
    for N, K in specs:
        A, W = padded_bf16(M, K, gen=g), padded_bf16(N, K, scale=0.05, gen=g)
        bias = 0.1 * torch.randn(N, device=DEV, generator=g)
        out = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        probs.append(make_problem(A, W, out, M, N, K, bias=bias))
        outs.append(out)
        wants.append(torch.nn.functional.elu(A.float()[:, :K] @ W.float()[:, :K].t() + bias))
        probs[-1]._keep = (A, W, bias)
    gemm(probs, _lib.EPI_BIAS_ELU_BF16, 128, cg)
    for i, (o, w) in enumerate(zip(outs, wants)):
        ok, msg = report(f"grouped[{i}]", o.float(), w, atol=2e-3, rtol=1e-2)
        assert ok, msg


def test_rejects_bad_arguments():
    A = padded_bf16(128, 64)
    out = torch.zeros(128, 128, device=DEV)
    with pytest.raises(_lib.HgError):
        gemm([make_problem(A, A, out, 128, 128, 64)], _lib.EPI_STORE_F32, 64)          # bad tile width
    with pytest.raises(_lib.HgError):
        gemm([make_problem(A, A, out, 128, 128, 64, k_splits=2)], _lib.EPI_STORE_F32, 128)  # split-K without atomics
