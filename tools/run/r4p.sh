mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_roc_decode -s 8 -c 1 -o gpurun_out/r5a_dec -f python tools/probe.py --n 2e8 --zipf 1.0 --ef 0 --reps 2 > gpurun_out/r5a_ncu.log 2>&1
tail -2 gpurun_out/r5a_ncu.log
