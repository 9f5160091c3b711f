set -x
mkdir -p gpurun_out
bash tools/gpu_suite.sh > gpurun_out/suite.log 2>&1; echo "suite rc $?" >> gpurun_out/suite.log
cat gpurun_out/suite.log | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
bash tools/sweep_envs.sh > gpurun_out/sweep.log 2>&1; tail -9 gpurun_out/sweep.log
