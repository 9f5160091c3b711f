set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh tests/test_gemm_gpu.py tests/test_learner_kernels_gpu.py tests/test_ppo_gpu.py > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -5
echo "== m-major"; HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
echo "== problem-major"; HG_TILE_ORDER=p HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
echo "== m-major"; HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
echo "== problem-major"; HG_TILE_ORDER=p HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
python bench.py --no-cpu-baseline --steps 20 > gpurun_out/bench_mmajor.json 2>/dev/null; python -c "
import json; b=json.load(open('gpurun_out/bench_mmajor.json')); print(b['ms_per_step'], b['roofline']['by_kind'])"
