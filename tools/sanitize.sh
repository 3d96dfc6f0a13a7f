# compute-sanitizer evidence: racecheck (the WarpMem unions alias kinematics scratch, the row
# matrix, the limit list and the obstacle staging in shared memory) and memcheck over every kernel of every env kind.
# usage: bash tools/sanitize.sh <tag>      -> gpurun_out/<tag>_racecheck.log, gpurun_out/<tag>_memcheck.log
cd ${GRAFT_REPO_ROOT:-.}
TAG=${1:-r2}
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 python tools/sanitize_run.py 64 3 > gpurun_out/${TAG}_racecheck.log 2>&1
tail -5 gpurun_out/${TAG}_racecheck.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 50 python tools/sanitize_run.py 64 3 > gpurun_out/${TAG}_memcheck.log 2>&1
tail -5 gpurun_out/${TAG}_memcheck.log
