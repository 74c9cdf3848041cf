python bench.py --no-e2e --no-cpu-baseline --no-ef --no-wt --no-sharded --no-accessors 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k in ('s05','control'):
    print(k, d['configs'][k]['ms_per_step'], d['configs'][k]['kernel_ms'])
print(d['clocks'])
"
