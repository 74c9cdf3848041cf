mkdir -p gpurun_out
for n in 1 2 3 4; do echo "ctas/sm $n"; IDC_EF_ENC_CTAS_PER_SM=$n python tools/ef_probe.py 1e9 0 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['encode_ms'],3))"; done
