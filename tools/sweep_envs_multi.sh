#!/bin/bash
# BASELINE.json configs[4] at N > 1: envs-per-GPU sweep under torchrun (one rank per GPU, NCCL), device-timed value + update-GEMM
# TFLOP/s + GAE / env-step GB/s per size; one JSON line each in gpurun_out/env_sweep_n$N.jsonl.
# Usage (on a box with N GPUs): bash tools/sweep_envs_multi.sh N [sizes...]
N=$1; shift
mkdir -p gpurun_out
sizes="$@"
[ -z "$sizes" ] && sizes="1024 2048 4096 8192 16384 32768 65536"
: > gpurun_out/env_sweep_n$N.jsonl
port=29560
for e in $sizes; do
  steps=10; [ $e -ge 16384 ] && steps=4
  port=$((port + 1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --envs $e --steps $steps --warmup 3 --no-e2e > gpurun_out/sweep_n${N}_$e.log 2> gpurun_out/sweep_n${N}_$e.err
  rc=$?
  grep '^{' gpurun_out/sweep_n${N}_$e.log | tail -1 >> gpurun_out/env_sweep_n$N.jsonl || echo "{\"envs_per_gpu\": $e, \"failed\": $rc}" >> gpurun_out/env_sweep_n$N.jsonl
  echo "N $N envs $e exit $rc"
done
python tools/sweep_summary.py gpurun_out/env_sweep_n$N.jsonl
