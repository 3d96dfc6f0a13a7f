# round-2 evidence run on one B200 (everything lands in gpurun_out/, the keepers are copied to profiles/ by hand): usage bash tools/final_run.sh [tag]
set -x
cd $GRAFT_REPO_ROOT
T=${1:-r3z}
python -m pytest tests -m gpu -q > gpurun_out/${T}_gputests.log 2>&1; tail -3 gpurun_out/${T}_gputests.log
( time python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -4 gpurun_out/${T}_bench_default.err
python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
for e in stepper monkey cassie child mike walker2d crab2d; do python bench.py --env $e --no-cpu-baseline > gpurun_out/${T}_bench_$e.json 2>gpurun_out/${T}_bench_$e.err; done
python bench.py --env cassie --envs 8192 --no-cpu-baseline > gpurun_out/${T}_bench_cassie8192.json 2>/dev/null
python bench.py --actions pd --no-cpu-baseline --no-also > gpurun_out/${T}_bench_pd.json 2>/dev/null
python bench.py --self-collision 0 --no-cpu-baseline --no-also > gpurun_out/${T}_bench_sc0.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-also > gpurun_out/${T}_ncu_bench.log 2>&1
bash tools/r3_prof.sh ${T}_custom custom 'k_step_walker3d_custom$'
bash tools/r3_prof.sh ${T}_stepper stepper 'k_step_walker3d_stepper$'
bash tools/r3_prof.sh ${T}_monkey monkey 'k_step_monkey3d_custom$'
bash tools/r3_prof.sh ${T}_cassie cassie 'k_step_cassie$' --envs 8192
ls -la gpurun_out | grep ${T}
for f in gpurun_out/${T}_bench_*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d.get('value'), d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))
except Exception as e: print('$f', 'ERR', e)"; done
