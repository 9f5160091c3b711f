import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_gemm import *
for bn in (128, 256):
    show(f"fwd L1 actor only bn{bn}", *fwd([(512, 705)], bn))
    show(f"fwd L2 actor only bn{bn}", *fwd([(256, 512)], bn))
    show(f"fwd K=2048 N=512 bn{bn}", *fwd([(512, 2048)], bn))
    show(f"fwd K=8192 N=512 bn{bn}", *fwd([(512, 8192)], bn, m=16384))
