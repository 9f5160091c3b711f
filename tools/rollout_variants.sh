#!/bin/bash
# Rollout-phase GEMM tile variants (tools/phase_times.py, 4096 envs): tile width x CTA pairing of the first two layers.
mkdir -p gpurun_out
for v in "128 1" "128 2" "256 1" "256 2"; do
  set -- $v
  echo "== HG_ROLLOUT_BN=$1 HG_ROLLOUT_CG=$2"
  HG_ROLLOUT_BN=$1 HG_ROLLOUT_CG=$2 HG_ITERS=6 timeout 300 python tools/phase_times.py 2>&1 | tail -1
done
