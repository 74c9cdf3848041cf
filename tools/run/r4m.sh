mkdir -p gpurun_out
IDC_DEC_DEBUG=gpurun_out/dec_dbg.bin python tools/adv_probe.py 2>&1 | tail -8
