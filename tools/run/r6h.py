"""Launch list of the short-unit ROC kernels: C2-sized lists (one unit per warp) and K = 64 graph rows (one row per
thread / per warp). Run under `ncu --metrics gpu__time_duration.sum`."""
import numpy as np
import torch

from vector_db_id_compression_b200.capi import Context

rng = np.random.default_rng(2)
lab = rng.integers(0, 1024, size=1_000_000)
order = np.argsort(lab, kind="stable").astype(np.int64)
off = np.zeros(1025, np.uint64)
off[1:] = np.cumsum(np.bincount(lab, minlength=1024))
ctx = Context(0)
b = ctx.roc_encode(off, torch.from_numpy(order).cuda(), sorted_ids=True)
b.decode(device="cuda")
N, K = 100_000, 64
rows = torch.stack([torch.randperm(N, device="cuda")[:K] for _ in range(64)]).repeat(N // 64 + 1, 1)[:N].int().contiguous()
g = ctx.roc_encode_rows(rows)
g.decode_rows(device="cuda")
g.decode_rows(np.arange(8, dtype=np.int32))
