mkdir -p gpurun_out
IDC_TRACE_HOST=1 python tools/rows_trace.py 2> gpurun_out/r2k_rows_trace.txt
cat gpurun_out/r2k_rows_trace.txt
