set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/r1v_gputests.log 2>&1; tail -3 gpurun_out/r1v_gputests.log
python bench.py > gpurun_out/r1v_bench_custom.json 2> gpurun_out/r1v_bench_custom.err
python bench.py --impl reference > gpurun_out/r1v_bench_ref.json 2> gpurun_out/r1v_bench_ref.err
for e in stepper monkey cassie child mike walker2d crab2d; do python bench.py --env $e --no-cpu-baseline > gpurun_out/r1v_bench_$e.json 2>gpurun_out/r1v_bench_$e.err; done
python bench.py --env cassie --envs 8192 --no-cpu-baseline > gpurun_out/r1v_bench_cassie8192.json 2>/dev/null
python bench.py --actions pd --no-cpu-baseline > gpurun_out/r1v_bench_pd.json 2>/dev/null
python bench.py --self-collision 0 --no-cpu-baseline > gpurun_out/r1v_bench_sc0.json 2>/dev/null
python tools/e2e_direct_ab.py > gpurun_out/r1v_e2e_direct_ab.txt 2>&1
python tools/e2e_breakdown.py > gpurun_out/r1v_e2e_breakdown.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 200 --csv --log-file gpurun_out/r1v_launches.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/r1v_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_walker3d_custom -s 12 -c 2 -f -o gpurun_out/prof_r1v python bench.py --steps 5 --warmup 10 --no-cpu-baseline > gpurun_out/r1v_ncu_full.log 2>&1
ls -la gpurun_out | tail -30
for f in gpurun_out/r1v_bench_*.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', d.get('value'), d.get('e2e',{}).get('value'))"; done
