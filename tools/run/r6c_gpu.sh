#!/bin/bash
# one GPU call: full parity suite, accessor bench, C3 probe with the thread-per-row ROC decoder and with the lane-group
# kernel (IDC_ROC_ROWS_GROUP=1), memcheck of the new test
mkdir -p gpurun_out
timeout 150 python -m pytest tests -q -m gpu > gpurun_out/r6c_pytest.log 2>&1; tail -4 gpurun_out/r6c_pytest.log
timeout 60 tools/accessor_bench > gpurun_out/r6c_acc.json 2> gpurun_out/r6c_acc.err; cat gpurun_out/r6c_acc.json
timeout 60 python tools/graph_probe.py 1e6 > gpurun_out/r6c_c3_small.txt 2>&1; grep ROC gpurun_out/r6c_c3_small.txt
IDC_ROC_ROWS_GROUP=1 timeout 60 python tools/graph_probe.py 1e6 > gpurun_out/r6c_c3_group.txt 2>&1; grep ROC gpurun_out/r6c_c3_group.txt
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "test_roc_rows_thread_per_row_decoder and (64-1000000 or 1-9)" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -6 | tee gpurun_out/r6c_memcheck.txt
