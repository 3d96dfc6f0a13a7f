# round-2 evidence run on one B200 (everything lands in gpurun_out/, the keepers are copied to profiles/ by hand)
set -x
cd $GRAFT_REPO_ROOT
T=r2z
python -m pytest tests -m gpu -q > gpurun_out/${T}_gputests.log 2>&1; tail -3 gpurun_out/${T}_gputests.log
( time python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; tail -4 gpurun_out/${T}_bench_default.err
python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
for e in stepper monkey cassie child mike walker2d crab2d; do python bench.py --env $e --no-cpu-baseline > gpurun_out/${T}_bench_$e.json 2>gpurun_out/${T}_bench_$e.err; done
python bench.py --self-collision 0 --no-cpu-baseline --no-also > gpurun_out/${T}_bench_sc0.json 2>/dev/null
python bench.py --env cassie --self-collision 0 --no-cpu-baseline > gpurun_out/${T}_bench_cassie_sc0.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-also > gpurun_out/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_walker3d_custom -s 12 -c 2 -f -o gpurun_out/prof_${T} python bench.py --steps 5 --warmup 10 --no-cpu-baseline --no-also > gpurun_out/${T}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_cassie -s 6 -c 1 -f -o gpurun_out/prof_${T}_cassie python bench.py --env cassie --envs 8192 --steps 3 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_ncu_cassie.log 2>&1
ls -la gpurun_out | grep ${T}
for f in gpurun_out/${T}_bench_*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d.get('value'), d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'))
except Exception as e: print('$f', 'ERR', e)"; done
