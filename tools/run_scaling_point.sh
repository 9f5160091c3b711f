#!/bin/bash
# One point of the weak-scaling curve exactly as the driver launches it: bash tools/run_scaling_point.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc $?"; grep "^{" gpurun_out/bench_n$N.json | cut -c1-260
