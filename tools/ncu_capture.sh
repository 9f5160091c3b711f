#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box AFTER the same program exited 0 without ncu):
#   1. launch list of one PPO iteration (gpu__time_duration per launch; cold-cache, serialised: shares only)
#   2. `--set full` of one minibatch's 8 grouped-GEMM launches (the rollout issues 60 layer-1 launches + 3 for the last
#      values before them), of one env-step launch, of one rollout step's layer-1 + tail launches and of one loss-head launch
# Reports land in gpurun_out/ncu_*.ncu-rep; summaries (tools/ncu_summary.py, tools/launch_summary.py) next to them.
mkdir -p gpurun_out
set -x
python tools/prof_iter.py > gpurun_out/prof_iter_plain.log 2>&1 || { echo "prof_iter failed without ncu"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launch_list_iteration.csv \
    python tools/prof_iter.py > gpurun_out/ncu_launch_list.log 2>&1
python tools/launch_summary.py gpurun_out/launch_list_iteration.csv > gpurun_out/launch_list_iteration_summary.txt 2>&1
# the same command the driver times, under the same single-metric pass (kernel SHARES of the step; absolutes are cold-cache)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_plain_steps2.json 2>/dev/null || { echo "bench failed without ncu"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launch_list_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_list_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launch_list_bench_steps2.csv > gpurun_out/launch_list_bench_steps2_summary.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel --launch-skip 63 -c 8 \
    -f -o gpurun_out/ncu_update_gemms python tools/prof_iter.py > gpurun_out/ncu_gemms.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_update_gemms.ncu-rep > gpurun_out/ncu_update_gemms.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:env_step_kernel --launch-skip 30 -c 1 \
    -f -o gpurun_out/ncu_env_step python tools/prof_iter.py > gpurun_out/ncu_env.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_env_step.ncu-rep 0 40 > gpurun_out/ncu_env_step.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:rollout_tail|gemm_kernel" --launch-skip 20 -c 2 \
    -f -o gpurun_out/ncu_rollout_step python tools/prof_iter.py > gpurun_out/ncu_rollout.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_rollout_step.ncu-rep > gpurun_out/ncu_rollout_step.txt 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ppo_head_loss --launch-skip 2 -c 1 \
    -f -o gpurun_out/ncu_loss_head python tools/prof_iter.py > gpurun_out/ncu_loss.log 2>&1
python tools/ncu_summary.py gpurun_out/ncu_loss_head.ncu-rep 0 25 > gpurun_out/ncu_loss_head.txt 2>&1
ls -la gpurun_out/*.ncu-rep
tail -12 gpurun_out/ncu_update_gemms.txt
