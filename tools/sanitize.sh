#!/bin/bash
# compute-sanitizer over the hot path (the reference has no race / memory checking at all).
#   tools/sanitize.sh memcheck|racecheck|initcheck|synccheck
# ONE tool per gpurun call (B200_PROFILING.md).  Workload: tools/sanitize_workload.py -- smoke()-sized learner check, one
# 256-env x 4-step PPO iteration in the history-ring format with resets (env kernel, ring gather / rollover, rollout GEMMs +
# tail kernel, GAE, loss head, dW / dX GEMMs, optimiser), and the stacked format's env step (in-place history shift).
# The races SURVEY 5 names are all in there: history shift in place, episode_sums read-modify-write, observation buffers
# aliasing the transition record.  Summary -> gpurun_out/sanitize_<tool>.txt (copy into profiles/).
set -u
tool=${1:-memcheck}
out=gpurun_out/sanitize_${tool}.txt
mkdir -p gpurun_out
python tools/sanitize_workload.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed: not sanitizing"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
compute-sanitizer --tool ${tool} --print-limit 20 --error-exitcode 3 python tools/sanitize_workload.py > ${out}.full 2>&1
rc=$?
{ echo "compute-sanitizer --tool ${tool} python tools/sanitize_workload.py  -> exit ${rc}"; grep -E "=========|workload" ${out}.full | tail -40; } > ${out}
cat ${out}
exit ${rc}
