mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r5c_bench2.json 2> gpurun_out/r5c_bench2.err; tail -c 300 gpurun_out/r5c_bench2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r5c_bench2.json').read().strip().splitlines()[-1])
e=d['e2e']; print('n_gpus',d['n_gpus'],'value',d['value']/1e9,'e2e',e['value']/1e9,e['ms_per_step'],e.get('roundtrip_all_lists_ok'),'pipelined',e.get('pipelined',{}).get('ms_per_step'))
print(json.dumps(d['sharded'])[:700])
P
