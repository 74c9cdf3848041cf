python bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value',d['value']/1e9,'e2e',d['e2e']['ms_per_step'])
for k in ('s05','control'):
    print(k, d['configs'][k]['ms_per_step'], d['configs'][k]['kernel_ms'])
"
