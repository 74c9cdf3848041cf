#!/bin/bash
# one GPU call: full parity suite, accessor bench with / without the mailbox path, memcheck of the mailbox test
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/r6a_pytest.log 2>&1; tail -4 gpurun_out/r6a_pytest.log
timeout 60 tools/accessor_bench > gpurun_out/r6a_acc_mailbox.json 2> gpurun_out/r6a_acc.err; cat gpurun_out/r6a_acc_mailbox.json
IDC_NO_MAILBOX=1 timeout 60 tools/accessor_bench > gpurun_out/r6a_acc_plain.json 2>> gpurun_out/r6a_acc.err; cat gpurun_out/r6a_acc_plain.json
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -k "test_small_host_calls" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -6 | tee gpurun_out/r6a_memcheck.txt
