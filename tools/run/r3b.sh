mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r3b_bench_n2.json 2> gpurun_out/r3b_bench_n2.err
tail -3 gpurun_out/r3b_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3b_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
e=d['e2e']; print({k:e[k] for k in e if k not in ('path',)})
print(json.dumps(d['sharded'],indent=0)[:2500])
PY
