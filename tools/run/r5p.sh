mkdir -p gpurun_out
K='test_roc_host_download_by_milestones or test_roc_host_buffers_pipeline or test_roc_adversarial or test_roc_graph_rows or test_roc_translate'
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error" | tail -4 > gpurun_out/r5p_memcheck.txt
cat gpurun_out/r5p_memcheck.txt
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_wavelet.py -x -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error" | tail -3 | tee -a gpurun_out/r5p_memcheck.txt
