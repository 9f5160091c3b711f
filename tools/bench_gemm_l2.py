import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_gemm as bg
import torch
for flush in (True, False):
    if not flush:
        bg.flush = torch.empty(8, dtype=torch.uint8, device=bg.DEV)   # no L2 flush: operands stay L2-resident
    for m in (8192, 16384, 61440):
        for bn in (256,):
            bg.show(f"flush={flush} M={m} fwd L1 actor only bn{bn}", *bg.fwd([(512, 705)], bn, m))
            bg.show(f"flush={flush} M={m} fwd N=512 K=2048 bn{bn}", *bg.fwd([(512, 2048)], bn, m))
