"""Epilogue-bound GEMM shapes for one-kernel ncu captures: usage prof_epi.py [fwd|dx] [bn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_gemm as bg
import bench_gemm_dx as bd
kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 128
shapes = [(256, 512), (256, 768)]
if kind == "fwd":
    bg.show(f"dx-shaped fwd epilogue bn{bn}", *bd.fwd_like(shapes, bn))
else:
    bg.show(f"dx colsum bn{bn}", *bd.dx(shapes, bn, True))
