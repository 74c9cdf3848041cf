mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -k "host" 2>&1 | grep -v "^$" | tail -3
IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "roc_encode" > gpurun_out/r4d_e2e.txt; tail -32 gpurun_out/r4d_e2e.txt
