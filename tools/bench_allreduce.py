"""All-reduce of the gradient arena (1,033,380 floats) under torchrun: isolated latency and the cost inside a kernel chain."""
import os, sys, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
g = torch.randn(1033380, device="cuda")
a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_ar = timed(lambda: dist.all_reduce(g))
t_mm = timed(lambda: torch.mm(a, a))
def chain():
    torch.mm(a, a); dist.all_reduce(g); torch.mm(a, a)
t_chain = timed(chain)
t = torch.tensor([t_ar, t_mm, t_chain], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: all_reduce(4.13 MB) {t[0]:.1f} us | mm {t[1]:.1f} us | mm+all_reduce+mm {t[2]:.1f} us -> all-reduce in a chain costs {t[2] - 2 * t[1]:.1f} us")
dist.destroy_process_group()
