IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep "roc_encode\|class \|encode(host)" | tail -18
