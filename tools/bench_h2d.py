"""Raw pinned host -> device bandwidth of this box (what bounds bench.py's e2e figure: 264 MB of Isaac-Gym-shaped state per
iteration at 4096 envs x 60 steps), as one large copy and as the 4.4 MB per-step copies the env streams."""
import torch

dev = "cuda:0"
for mb, reps in ((256, 10), (4.4, 240)):
    n = int(mb * 1e6 / 4)
    src = torch.empty(n, dtype=torch.float32).pin_memory()
    dst = torch.empty(n, dtype=torch.float32, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        dst.copy_(src, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"H2D pinned {mb} MB x {reps}: {n * 4 * reps / ms / 1e6:.1f} GB/s ({ms / reps * 1e3:.0f} us per copy)")
