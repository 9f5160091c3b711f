import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_gemm as bg
import torch
from gpu_util import make_problem
from humanoid_gym_with_comments_b200 import _lib

def dx(specs, bn, colsum=True, m=bg.M):
    probs, keep, fl = [], [], 0
    for NL, KL in specs:
        dZ, W = bg.bf(m, NL, 1e-3), bg.bf(NL, KL, 0.05)
        H = bg.bf(m, KL); out = torch.empty(m, KL, dtype=torch.bfloat16, device=bg.DEV); cs = torch.zeros(KL, device=bg.DEV)
        probs.append(make_problem(dZ, W, out, m, KL, NL, 0, 1, act=H, colsum=cs if colsum else None)); keep.append((dZ, W, H, out, cs)); fl += 2.0 * m * NL * KL
    ms = bg.timeit(lambda: bg.gemm_nosync(probs, _lib.EPI_DELU_BF16, bn))
    return ms, fl

def fwd_like(specs, bn, m=bg.M):   # same shapes through the forward epilogue (no act read, no colsum)
    probs, keep, fl = [], [], 0
    for NL, KL in specs:
        dZ, W = bg.bf(m, NL, 1e-3), bg.bf(NL, KL, 0.05)
        out = torch.empty(m, KL, dtype=torch.bfloat16, device=bg.DEV); bias = torch.zeros(KL, device=bg.DEV)
        probs.append(make_problem(dZ, W, out, m, KL, NL, 0, 1, bias=bias)); keep.append((dZ, W, out, bias)); fl += 2.0 * m * NL * KL
    ms = bg.timeit(lambda: bg.gemm_nosync(probs, _lib.EPI_BIAS_ELU_BF16, bn))
    return ms, fl

if __name__ == '__main__':
  for bn in (128, 256):
    bg.show(f"dx L2 group bn{bn} colsum", *dx([(256, 512), (256, 768)], bn, True))
    bg.show(f"dx L2 group bn{bn} NO colsum", *dx([(256, 512), (256, 768)], bn, False))
    bg.show(f"dx-shaped L2 group bn{bn} fwd epilogue", *fwd_like([(256, 512), (256, 768)], bn))
