#!/bin/bash
# compute-sanitizer over a reduced GPU parity run (memcheck, then racecheck + synccheck on the ROC / EF kernels)
set -u
K='test_roc_random_lists_vs_oracle or test_roc_graph_rows or test_ef_lists_vs_oracle or test_ef_graph_rows or test_roc_adversarial or test_roc_translate or test_packed_bits or test_ef_long_list'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -6
done
