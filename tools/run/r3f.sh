mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 8 python -m pytest tests/test_gpu_parity.py -x -q -k "test_ef_lists_vs_oracle or test_ef_bulk_copy_alignment or test_ef_long_list or test_ef_graph_rows" > gpurun_out/r3f_memcheck.txt 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r3f_memcheck.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -x -q -k "test_ef_lists_vs_oracle or test_ef_bulk_copy_alignment or test_ef_graph_rows" > gpurun_out/r3f_racecheck.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported" gpurun_out/r3f_racecheck.txt | tail -5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ef_ -s 12 -c 6 -o gpurun_out/r3f_ef -f python tools/ef_probe.py 1e9 1.0 > gpurun_out/r3f_ncu.log 2>&1
tail -1 gpurun_out/r3f_ncu.log
