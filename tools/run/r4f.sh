mkdir -p gpurun_out
IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "job " > gpurun_out/r4f_e2e.txt; tail -34 gpurun_out/r4f_e2e.txt
