"""Wavelet-tree flavour on one GPU: build, whole-index decode and random get_single_id timings (CUDA events of the
codec context). python tools/wt_probe.py [n_ids] [nlist]"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vector_db_id_compression_b200.capi import Context  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nlist = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
REPS = int(os.environ.get("WT_PROBE_REPS", "3"))  # 1 under ncu: every kernel once
g = torch.Generator(device="cuda").manual_seed(7)
lab = torch.randint(0, nlist, (n,), device="cuda", generator=g)
ids = torch.argsort(lab, stable=True)  # a random partition of [0, n), ascending per list
sizes = torch.bincount(lab, minlength=nlist).cpu().numpy()
offsets = np.zeros(nlist + 1, np.uint64)
offsets[1:] = np.cumsum(sizes)
del lab
ctx = Context(0)
ctx.set_timing(True)
res = dict(n_ids=n, nlist=nlist)
for rep in range(REPS):
    blob = ctx.wt_encode(offsets, ids)
    bd = {}
    for name, ms in ctx.last_kernel_breakdown():
        bd[name] = bd.get(name, 0.0) + ms
    res["encode_ms"] = ctx.last_kernel_ms()
    res["encode_breakdown_ms"] = bd
    if rep < REPS - 1:
        blob.free()
res.update(levels=blob.levels, bits_bytes=blob.bits_bytes, aux_bytes=blob.aux_bytes,
           bits_per_id=8.0 * (blob.bits_bytes + blob.aux_bytes) / n)
# build: per level 4 B read (bits) + 4 B read + 4 B written (partition) per id, except the last level
res["encode_algorithmic_GBs"] = (n * 4.0 * (3 * blob.levels - 2) + n * 8.0) / res["encode_ms"] / 1e6
for rep in range(min(REPS, 2)):
    dec, _ = blob.decode(device="cuda")
    res["decode_all_ms"] = ctx.last_kernel_ms()
    bd = {}
    for name, ms in ctx.last_kernel_breakdown():
        bd[name] = bd.get(name, 0.0) + ms
    res["decode_all_breakdown_ms"] = bd
assert torch.equal(dec, ids), "decode mismatch"
res["decode_all_Mids_s"] = n / res["decode_all_ms"] / 1e3
nq = 10_000_000
ql = torch.randint(0, nlist, (nq,), device="cuda", generator=g)
sz = torch.from_numpy(sizes).cuda()[ql]
ql, sz = ql[sz > 0], sz[sz > 0]
qo = (torch.rand(ql.numel(), device="cuda", generator=g) * sz).long()
qo = torch.minimum(qo, sz - 1)
for rep in range(min(REPS, 2)):
    got = blob.select(ql, qo, device="cuda")
    res["select_ms"] = ctx.last_kernel_ms()
want = ids[torch.from_numpy(offsets[:-1].astype(np.int64)).cuda()[ql] + qo]
assert torch.equal(got, want), "select mismatch"
res["select_queries"] = int(ql.numel())
res["select_Mq_s"] = ql.numel() / res["select_ms"] / 1e3
print(json.dumps(res))
