# same-box A/B of the working library against saved variants under ab/: usage  bash tools/r2_ab.sh <tag> "<envs>" [label=lib ...]
cd $GRAFT_REPO_ROOT
TAG=$1; ENVS=$2; shift; shift
for e in $ENVS; do ENVK=$e STEPS=${STEPS:-400} bash tools/abbench.sh "$@" cur=mocca_envs_b200/libmocca_b200.so; done 2>&1 | grep -v "^+" | tee gpurun_out/${TAG}_ab.txt
