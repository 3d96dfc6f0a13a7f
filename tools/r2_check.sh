# round-2 sanity run on the GPU box: GPU tests + the default bench line
set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; tail -5 gpurun_out/r2a_gputests.log
python bench.py --no-cpu-baseline > gpurun_out/r2a_bench_custom.json 2> gpurun_out/r2a_bench_custom.err; tail -c 600 gpurun_out/r2a_bench_custom.json
