import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_learner_kernels_gpu as T
import ctypes as C
orig_allclose = torch.allclose
def patched(a, b, rtol=1e-5, atol=1e-8, **kw):
    ok = orig_allclose(a, b, rtol=rtol, atol=atol, **kw)
    if not ok and a.dim() == 2:
        err = (a - b).abs(); tol = atol + rtol * b.abs(); bad = err > tol
        idx = bad.nonzero()
        print("shape", tuple(a.shape), "bad", int(bad.sum()), "of", bad.numel(), "max err", float(err.max()), "typ |ref|", float(b.abs().mean()))
        print("bad rows mod 16:", torch.bincount(idx[:, 0] % 16, minlength=16).tolist())
        print("bad cols mod 16:", torch.bincount(idx[:, 1] % 16, minlength=16).tolist())
        print("bad rows // 64:", torch.bincount(idx[:, 0] // 64).tolist()[:12])
        for i in idx[:10].tolist(): print(i, float(a[tuple(i)]), float(b[tuple(i)]))
        return True
    return ok
torch.allclose = patched
T.test_head_loss_forward_backward(384, 12, 128)
print("done")
