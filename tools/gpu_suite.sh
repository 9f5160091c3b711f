#!/bin/bash
# Runs every -m gpu test file in its own process (a faulting kernel poisons its CUDA context) with a timeout,
# logs under gpurun_out/.  Usage (on the GPU box): bash tools/gpu_suite.sh [files...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
files="$@"
[ -z "$files" ] && files="tests/test_gemm_gpu.py tests/test_gae_gpu.py tests/test_env_gpu.py tests/test_learner_kernels_gpu.py tests/test_pre_physics.py tests/test_ppo_gpu.py"
rc_all=0
for f in $files; do
  log=gpurun_out/$(basename $f .py).log
  timeout 600 python -m pytest $f -q -m gpu -s -p no:cacheprovider > $log 2>&1
  rc=$?
  echo "== $f exit $rc: $(tail -1 $log)"
  grep -E "^(FAILED|ERROR)" $log | head -40
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
