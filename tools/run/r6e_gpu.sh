#!/bin/bash
# one GPU call: full parity suite + the full default bench line of the final build
mkdir -p gpurun_out
timeout 150 python -m pytest tests -q -m gpu > gpurun_out/r6e_pytest.log 2>&1; tail -5 gpurun_out/r6e_pytest.log
timeout 170 python bench.py > gpurun_out/r6e_bench.json 2> gpurun_out/r6e_bench.err; tail -c 600 gpurun_out/r6e_bench.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r6e_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"]["bit_exact"])
    for k in ("c1", "c2", "c3", "c4"):
        print(k, json.dumps(d["configs"][k])[:900])
    print(json.dumps(d["accessors"])[:1500])
    print(d["leg_seconds"])
except Exception as e:
    print("bench line unreadable:", e)
P
