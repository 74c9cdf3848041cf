mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -1
(nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active,temperature.gpu --format=csv -lms 500 > gpurun_out/r5o_smi.csv 2>&1 &) 
python bench.py 2>/dev/null > gpurun_out/r5o_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r5o_bench.json').read().strip().splitlines()[-1])
print('value',d['value']/1e9,'e2e',d['e2e']['ms_per_step'])
for k in ('s05','control'):
    print(k, d['configs'][k]['ms_per_step'], d['configs'][k]['kernel_ms']['k_roc_encode'], d['configs'][k]['kernel_ms']['k_roc_decode'])
print(d['leg_seconds'])
"
pkill -x nvidia-smi || true
awk -F, 'NR>1{print $2,$4,$5}' gpurun_out/r5o_smi.csv | sort | uniq -c | sort -rn | head -8
