python tools/adv_probe.py 2>&1 | tail -5
IDC_ROC_G=2 python tools/adv_probe.py 2>&1 | tail -5
IDC_ROC_G=8 python tools/adv_probe.py 2>&1 | tail -5
