#!/bin/bash
# diagnostic: where the first ROC row-encode call of a process spends its time; the fixed test
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -q -k "test_roc_rows_thread_per_row_decoder" 2>&1 | tail -3
IDC_TRACE_HOST=1 timeout 60 tools/accessor_bench > gpurun_out/r6d_acc.json 2> gpurun_out/r6d_acc.err; grep -E "roc_encode_rows" gpurun_out/r6d_acc.err | head -20
timeout 60 python - <<'P' 2>&1 | tail -12
import time, numpy as np, torch
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W
dev = torch.device("cuda:0")
data, _ = W.nsg_like_graph(200000, 64, 3, dev)
ctx = Context(0)
torch.cuda.synchronize()
for k in range(4):
    t = time.perf_counter(); b = ctx.roc_encode_rows(data); ctx.synchronize(); print("encode_rows call", k, round(1e3 * (time.perf_counter() - t), 2), "ms"); b.free()
P
