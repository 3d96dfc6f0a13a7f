# same-box A/B of library variants: usage  bash tools/abbench.sh "<label>=<lib>[:SORT[:SELFCOL]]" ...   (env custom unless ENVK set)
ENVK=${ENVK:-custom}
for spec in "$@"; do
  label=${spec%%=*}; rest=${spec#*=}; IFS=: read lib se sc <<< "$rest"; se=${se:-1}; sc=${sc:-1}
  MB200_LIB=$lib MB200_SORT_EVERY=$se python bench.py --env $ENVK --self-collision $sc --steps ${STEPS:-600} --warmup 100 --no-cpu-baseline --no-also 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$label', '$ENVK', 'sc=$sc', 'value %.3fM'%(d['value']/1e6), 'ms %.4f'%d['ms_per_step'], 'e2e %.3fM'%(d['e2e']['value']/1e6), 'rows %.2f'%d['roofline']['rows_per_substep'], 'clk', d['clocks']['sm_mhz'])"
done
