mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 8 python -m pytest tests/test_gpu_parity.py -x -q -k "test_ef_lists_vs_oracle or test_ef_bulk_copy_alignment or test_ef_long_list or test_ef_graph_rows or test_roc_graph_rows" > gpurun_out/r2y_memcheck.txt 2>&1
grep -n "=========" gpurun_out/r2y_memcheck.txt | grep -v "Host Frame" | head -30
tail -4 gpurun_out/r2y_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -x -q -k "test_ef_lists_vs_oracle or test_ef_graph_rows" > gpurun_out/r2y_racecheck_ef.txt 2>&1
grep -n "=========" gpurun_out/r2y_racecheck_ef.txt | grep -v "Host Frame" | head -30
tail -3 gpurun_out/r2y_racecheck_ef.txt
