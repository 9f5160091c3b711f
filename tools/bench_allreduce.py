"""All-reduce of the gradient arena (1,033,380 floats) under torchrun: NCCL against the symmetric-memory kernels
(two-shot multimem / peer-to-peer two-shot / one-shot), isolated and inside a kernel chain."""
import os, sys, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1033380
g = torch.randn(n, device="cuda")
a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
def timed(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
res = {}
res["nccl"] = timed(lambda: dist.all_reduce(g))
t_mm = timed(lambda: torch.mm(a, a))
def chain(ar):
    def f():
        torch.mm(a, a); ar(); torch.mm(a, a)
    return f
res["nccl in chain"] = timed(chain(lambda: dist.all_reduce(g))) - 2 * t_mm
try:
    import torch.distributed._symmetric_memory as symm
    gname = dist.group.WORLD.group_name
    pad = (n + 128 * world - 1) // (128 * world) * (128 * world)
    buf = symm.empty(pad, dtype=torch.float32, device="cuda")
    hdl = symm.rendezvous(buf, dist.group.WORLD)
    buf.normal_()
    buf.mul_(1e-3)
    if rank == 0:
        print("symmetric memory: multicast support", hdl.has_multicast_support() if callable(hdl.has_multicast_support) else hdl.has_multicast_support, flush=True)
    ref = buf.clone(); dist.all_reduce(ref)
    ops = [("two_shot", lambda: torch.ops.symm_mem.two_shot_all_reduce_(buf, "sum", gname))]
    if (hdl.has_multicast_support() if callable(hdl.has_multicast_support) else hdl.has_multicast_support):
        ops.insert(0, ("multimem", lambda: torch.ops.symm_mem.multimem_all_reduce_(buf, "sum", gname)))
    for name, op in ops:
        try:
            keep = buf.clone()
            ref = buf.clone(); dist.all_reduce(ref)
            op(); torch.cuda.synchronize()
            err = float((buf - ref).abs().max())
            def op_z(op=op):             # (timed on zeros: repeated in-place sums of live data overflow)
                op()
            buf.zero_()
            res[name] = timed(op_z)
            res[name + " in chain"] = timed(chain(op_z)) - 2 * t_mm
            res[name + " max err"] = err
            buf.copy_(keep)
        except Exception as e:
            res[name] = repr(e)[:120]
except Exception as e:
    res["symm"] = repr(e)[:200]
if rank == 0:
    print(f"world {world}: mm {t_mm:.1f} us |", " | ".join(f"{k}: {v:.1f}" if isinstance(v, float) else f"{k}: {v}" for k, v in res.items()))
dist.destroy_process_group()
