for se in 0 1 2 4; do MB200_SORT_EVERY=$se python bench.py --steps 600 --warmup 100 --no-cpu-baseline > gpurun_out/sort_$se.json 2>gpurun_out/sort_$se.err; python - <<PY
import json
d=json.loads(open('gpurun_out/sort_$se.json').read().strip().splitlines()[-1])
print('sort_every=$se', 'value %.3fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'e2e %.3fM'%(d['e2e']['value']/1e6), 'launches', d['gpu_launches'])
PY
done
MB200_SORT_EVERY=0 python bench.py --env stepper --steps 400 --warmup 100 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stepper sort0 %.3fM'%(d['value']/1e6))"
MB200_SORT_EVERY=1 python bench.py --env stepper --steps 400 --warmup 100 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stepper sort1 %.3fM'%(d['value']/1e6))"
MB200_SORT_EVERY=0 python bench.py --env monkey --steps 400 --warmup 100 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('monkey sort0 %.3fM'%(d['value']/1e6))"
MB200_SORT_EVERY=1 python bench.py --env monkey --steps 400 --warmup 100 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('monkey sort1 %.3fM'%(d['value']/1e6))"
