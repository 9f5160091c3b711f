set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -8
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python tools/bench_gemm_dx.py 2>&1 | tail -6
HG_ITERS=6 python tools/phase_times.py 2>&1 | tail -1
bash tools/ncu_capture.sh > gpurun_out/ncu_capture.log 2>&1; tail -15 gpurun_out/ncu_capture.log
