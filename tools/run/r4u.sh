IDC_TRACE_HOST=1 python tools/c4_trace.py 2>&1 | tail -22
