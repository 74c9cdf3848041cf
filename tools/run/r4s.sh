bash tools/ab.sh "IDC_X=0" 2>&1 | tail -1
bash tools/ab.sh "IDC_X=0" --zipf-s 0 2>&1 | tail -1
