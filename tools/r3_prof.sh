# ncu --set full capture of one step kernel: usage bash tools/r3_prof.sh <tag> <bench env> <kernel regex> [extra bench args]
cd $GRAFT_REPO_ROOT
TAG=$1; ENVK=$2; KRE=$3; shift; shift; shift
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 12 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --env $ENVK --steps 5 --warmup 10 --no-cpu-baseline --no-also "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
