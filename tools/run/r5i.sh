for c in 16 32 64; do echo "chunks=$c"; IDC_UPLOAD_CHUNKS=$c timeout 600 python tools/e2e_probe.py 2>&1 | grep "encode(host)" | tail -2; done
