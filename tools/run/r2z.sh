mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_gpu_parity.py -x -q -k "test_roc_random_lists_vs_oracle or test_roc_graph_rows or test_roc_adversarial or test_roc_translate" > gpurun_out/r2z_racecheck_roc.txt 2>&1
grep -n "=========" gpurun_out/r2z_racecheck_roc.txt | grep -v "Host Frame" | head -40
tail -3 gpurun_out/r2z_racecheck_roc.txt
