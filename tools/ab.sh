#!/bin/bash
# usage: tools/ab.sh "<ENV=VAL ...>" [bench args]  -> one line: value, encode ms, decode ms
envs="$1"; shift
out=$(env $envs python bench.py --steps 2 --warmup 2 --no-e2e --no-ef --no-cpu-baseline "$@" 2>/dev/null)
python - "$envs" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]; ms = {r["kernel"]: r["kernel_ms"]}; ms.update({k: v["ms"] for k, v in r["other"].items()})
print(sys.argv[1], "| %.3f Gids/s step %.1f ms enc %.1f dec %.1f" % (d["value"] / 1e9, d["ms_per_step"], ms["k_roc_encode"], ms["k_roc_decode"]))
PY
