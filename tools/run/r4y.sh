for p in 8 12 16; do echo "parts=$p"; IDC_MS_PARTS=$p IDC_TRACE_HOST=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep "encode(host)\|copies done\|kernels done" | tail -3; done
