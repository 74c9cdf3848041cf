python tools/adv_probe.py 2>&1 | tail -4
IDC_DEC_NO_DEFER=1 python tools/adv_probe.py 2>&1 | tail -4
