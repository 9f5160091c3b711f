set -x
mkdir -p gpurun_out
echo "== default"; timeout 300 python tools/bench_gemm_dx.py 2>&1 | grep "colsum" | grep -v NO
echo "== direct act"; HG_LIB_PATH=$PWD/variants/libhgb200_directact.so timeout 300 python tools/bench_gemm_dx.py 2>&1 | grep "colsum" | grep -v NO
HG_LIB_PATH=$PWD/variants/libhgb200_directact.so timeout 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "delu" 2>&1 | tail -2
echo "== default"; HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
echo "== direct act"; HG_LIB_PATH=$PWD/variants/libhgb200_directact.so HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
echo "== default"; HG_ITERS=8 timeout 300 python tools/phase_times.py 2>&1 | tail -1
