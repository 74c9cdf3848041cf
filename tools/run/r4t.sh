mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r4t_bench.json 2> gpurun_out/r4t_bench.err; tail -c 400 gpurun_out/r4t_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r4t_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print('value',d['value']/1e9,'ms',d['ms_per_step'],'e2e',e['value']/1e9,e['ms_per_step'],'pipelined',e.get('pipelined',{}).get('ms_per_step'),'4B',e.get('id_bytes_4',{}).get('ms_per_step'))
print(d['roofline'])
print(d.get('leg_seconds'))
P
