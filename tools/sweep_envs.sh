#!/bin/bash
# BASELINE.json configs[4]: envs-per-GPU sweep at N=1.  Runs bench.py per size (device-timed value, grouped-GEMM
# TFLOP/s by kind, GAE / env-step GB/s), one JSON line each, collected in gpurun_out/env_sweep.jsonl.
# Usage (on the GPU box): bash tools/sweep_envs.sh [sizes...]
mkdir -p gpurun_out
sizes="$@"
[ -z "$sizes" ] && sizes="1024 2048 4096 8192 16384 32768 65536"
: > gpurun_out/env_sweep.jsonl
for e in $sizes; do
  steps=10; [ $e -ge 16384 ] && steps=4
  timeout 600 python bench.py --envs $e --steps $steps --warmup 3 --no-cpu-baseline --no-ref-cuda --no-e2e > gpurun_out/sweep_$e.log 2>&1
  rc=$?
  tail -1 gpurun_out/sweep_$e.log | grep '^{' >> gpurun_out/env_sweep.jsonl || echo "{\"envs_per_gpu\": $e, \"failed\": $rc}" >> gpurun_out/env_sweep.jsonl
  echo "envs $e exit $rc"
done
python tools/sweep_summary.py gpurun_out/env_sweep.jsonl
