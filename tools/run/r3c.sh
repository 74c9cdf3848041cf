mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "roc" 2>&1 | tail -2
timeout 600 python bench.py --no-e2e --no-ef --no-wt --no-sharded --no-configs --no-accessors --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['other']['k_roc_encode']['ms'])
PY
