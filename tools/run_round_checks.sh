set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -8
grep -E "passed|failed|Error|assert" gpurun_out/test_env_gpu.log | tail -12
python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
bash tools/rollout_variants.sh > gpurun_out/rollout_variants.log 2>&1
cat gpurun_out/rollout_variants.log
bash tools/sweep_envs.sh 4096 32768 > gpurun_out/sweep.log 2>&1; tail -4 gpurun_out/sweep.log
