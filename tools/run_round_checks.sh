set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh tests/test_env_gpu.py tests/test_ppo_gpu.py > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -4
HG_ITERS=6 python tools/phase_times.py 2>&1 | tail -1
bash tools/sweep_envs.sh 4096 16384 65536 > gpurun_out/sweep.log 2>&1; tail -5 gpurun_out/sweep.log
