#!/bin/bash
# What the round-end driver runs, in one call on a GPU box: the -m gpu suite (one process per file), smoke(), the bench
# (both arms), then the ncu evidence for profiles/ (tools/ncu_capture.sh).  Logs under gpurun_out/.
set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; echo "reference rc $?"
[ -n "$SKIP_NCU" ] || { bash tools/ncu_capture.sh > gpurun_out/ncu_capture.log 2>&1; tail -14 gpurun_out/ncu_capture.log; }
