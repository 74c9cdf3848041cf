python -m pytest tests/test_gpu_wavelet.py -x -q 2>&1 | tail -1
python tools/wt_probe.py 1e9 > gpurun_out/r5g_wt.json 2>/dev/null; python -c "import json; d=json.loads(open('gpurun_out/r5g_wt.json').read().strip().splitlines()[-1]); print(d['encode_ms'], d['encode_breakdown_ms'], d['decode_all_ms'], d['decode_all_breakdown_ms'], d['select_ms'])"
