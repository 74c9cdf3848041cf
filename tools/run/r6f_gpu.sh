#!/bin/bash
# the decode scan unrolled by four: the tests of the short-unit kernels, C2 kernel times
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -q -k "short_units or thread_per_row or random_lists or golden" 2>&1 | tail -3
timeout 60 python - <<'P' 2>&1 | tail -6
import numpy as np, torch, time
from vector_db_id_compression_b200.capi import Context
rng = np.random.default_rng(2)
lab = rng.integers(0, 1024, size=1_000_000)
order = np.argsort(lab, kind="stable").astype(np.int64)
off = np.zeros(1025, np.uint64); off[1:] = np.cumsum(np.bincount(lab, minlength=1024))
ids = torch.from_numpy(order).cuda()
ctx = Context(0); ctx.set_timing(True)
for rep in range(3):
    b = ctx.roc_encode(off, ids, sorted_ids=True); enc = ctx.last_kernel_breakdown()
    d, _ = b.decode(device="cuda"); dec = ctx.last_kernel_breakdown()
    ok = bool(torch.equal(torch.sort(d.view(-1))[0], torch.arange(1_000_000, device="cuda")))
    b.free()
print("C2 encode", [(k, round(v, 3)) for k, v in enc], "decode", [(k, round(v, 3)) for k, v in dec], "set ok", ok)
P
