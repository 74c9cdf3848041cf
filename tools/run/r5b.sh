bash tools/ab.sh "IDC_ROC_G8_MIN_N=32768" 2>&1 | tail -1
bash tools/ab.sh "IDC_CLASS_STREAMS=4" 2>&1 | tail -1
bash tools/ab.sh "IDC_CLASS_STREAMS=2" 2>&1 | tail -1
bash tools/ab.sh "IDC_ROC_G2_MAX_N=8192" 2>&1 | tail -1
