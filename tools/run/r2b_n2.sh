set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_pytest_sharded.log
cat gpurun_out/r2b_pytest_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -5 gpurun_out/r2b_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_n2.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','sharded','leg_seconds')}, indent=1))
PY
