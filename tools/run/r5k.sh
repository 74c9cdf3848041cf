python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -2
IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep "launch \|encode(host)" | tail -14
