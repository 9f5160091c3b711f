import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_gemm as bg
for bn in (128, 256):
    bg.show(f"ldalign={os.environ.get('HG_LD_ALIGN')} fwd L1 actor only bn{bn}", *bg.fwd([(512, 705)], bn))
    bg.show(f"ldalign={os.environ.get('HG_LD_ALIGN')} fwd L1 a+c+l bn{bn}", *bg.fwd([(512, 705), (768, 219), (128, 705)], bn))
    bg.show(f"ldalign={os.environ.get('HG_LD_ALIGN')} fwd K=704 N=512 bn{bn}", *bg.fwd([(512, 704)], bn))
    bg.show(f"ldalign={os.environ.get('HG_LD_ALIGN')} fwd K=768 N=512 bn{bn}", *bg.fwd([(512, 768)], bn))
