mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -k "milestones" 2>&1 | grep -v "^$" | tail -40
IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "roc_encode" > gpurun_out/r4b_e2e.txt; tail -24 gpurun_out/r4b_e2e.txt
