# ncu --set full capture of the Walker3DCustom step kernel: usage bash tools/r2_prof.sh <tag> [lib]
cd $GRAFT_REPO_ROOT
TAG=$1; LIB=${2:-mocca_envs_b200/libmocca_b200.so}
MB200_LIB=$LIB ncu --set full --clock-control none --import-source on -k regex:k_step_walker3d_custom -s 12 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 5 --warmup 10 --no-cpu-baseline --no-also > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/prof_${TAG}.ncu-rep
