#!/bin/bash
# compute-sanitizer over a reduced GPU parity run (memcheck, then racecheck + synccheck) of the ROC / EF / wavelet kernels
set -u
K='test_roc_random_lists_vs_oracle or test_roc_graph_rows or test_ef_lists_vs_oracle or test_ef_graph_rows or test_ef_bulk_copy_alignment or test_roc_adversarial or test_roc_translate or test_packed_bits or test_ef_long_list'
K2='test_ef_import_and_file_round_trip or test_ef_row_blob_file_round_trip or test_roc_row_blob_file_round_trip'
K3='test_wt_type1_rrr_blocks_vs_oracle and (7-5000 or 65-20481)'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -6
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_blob_files.py -x -q -k "$K2" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -4
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_wavelet.py -x -q -k "$K3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -4
done
