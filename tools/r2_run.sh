# round-2 GPU batch: usage  bash tools/r2_run.sh <tag> [ab specs...]
set -x
cd $GRAFT_REPO_ROOT
TAG=$1; shift
python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_gputests.log 2>&1; tail -15 gpurun_out/${TAG}_gputests.log
python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print('headline %.3fM e2e %.3fM'%(d['value']/1e6,d['e2e']['value']/1e6), d['episodes']['cap_overflows'])
for k,v in d.get('also',{}).items(): print(k,'%.3fM e2e %.3fM frac %.3f'%(v['value']/1e6,v['e2e']['value']/1e6,v['roofline']['frac']))
PY
if [ $# -gt 0 ]; then
  for e in custom stepper cassie; do ENVK=$e STEPS=400 bash tools/abbench.sh base=mocca_envs_b200/libmocca_b200.so "$@" base2=mocca_envs_b200/libmocca_b200.so; done 2>&1 | grep -v "^+" | tee gpurun_out/${TAG}_ab.txt
fi
