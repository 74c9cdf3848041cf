mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -k "milestones" 2>&1 | grep -v "^$" | tail -5
IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "roc_encode" > gpurun_out/r4c_e2e.txt; tail -12 gpurun_out/r4c_e2e.txt
IDC_MS_PARTS=4 IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "roc_encode" > gpurun_out/r4c_e2e4.txt; tail -6 gpurun_out/r4c_e2e4.txt
