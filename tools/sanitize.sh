# compute-sanitizer evidence: racecheck (the WarpMem unions alias kinematics scratch, the row
# matrix, the limit list and the obstacle staging in shared memory) and memcheck over every kernel of every env kind.
# usage: bash tools/sanitize.sh <tag> [only-this-VecEnv-class-substring]
#   -> gpurun_out/<tag>_racecheck.log (per code-location-pair analysis records), gpurun_out/<tag>_memcheck.log
cd ${GRAFT_REPO_ROOT:-.}
TAG=${1:-r2}; ONLY=$2
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 400 python tools/sanitize_run.py 64 3 $ONLY > gpurun_out/${TAG}_racecheck.log 2>&1
grep -c "Race reported" gpurun_out/${TAG}_racecheck.log; tail -3 gpurun_out/${TAG}_racecheck.log
if [ -z "$SKIP_MEMCHECK" ]; then
timeout 1500 compute-sanitizer --tool memcheck --print-limit 50 python tools/sanitize_run.py 64 3 $ONLY > gpurun_out/${TAG}_memcheck.log 2>&1
tail -3 gpurun_out/${TAG}_memcheck.log
fi
