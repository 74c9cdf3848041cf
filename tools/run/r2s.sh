mkdir -p gpurun_out
timeout 900 python bench.py --no-ef --no-wt --no-sharded --no-configs --no-accessors --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
tail -3 gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d['e2e'],indent=0))
PY
