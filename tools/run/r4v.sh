mkdir -p gpurun_out
K='test_roc_random_lists_vs_oracle or test_roc_graph_rows or test_roc_adversarial or test_roc_translate or test_roc_host_download_by_milestones or test_roc_full_size_units'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard" | tail -8
done > gpurun_out/r4v_sanitize.txt 2>&1
cat gpurun_out/r4v_sanitize.txt
