"""Synthetic workloads of BASELINE.json / SURVEY.md 8(d), generated on the GPU with torch.

C5 ("1 B synthetic IDs in 65 536 Zipf-length lists"): list lengths L_k ~ k^-s normalised to sum N
(largest-remainder rounding, min 1); the ids of list k are the values found at positions
[off_k, off_k + L_k) of a seeded random permutation of [0, N) -- a uniformly random partition of the
id space -- ascending inside each list (Faiss add order).
"""
from __future__ import annotations

import numpy as np


def zipf_sizes(n_total: int, nlist: int, s: float) -> np.ndarray:
    if s == 0.0:
        w = np.ones(nlist, dtype=np.float64)
    else:
        w = np.arange(1, nlist + 1, dtype=np.float64) ** (-s)
    raw = w / w.sum() * (n_total - nlist)  # min 1 each
    base = np.floor(raw).astype(np.int64)
    rem = n_total - nlist - int(base.sum())
    order = np.argsort(-(raw - base), kind="stable")
    base[order[:rem]] += 1
    return base + 1


def random_partition_lists(n_total: int, sizes: np.ndarray, seed: int, device):
    """-> (offsets uint64 numpy [nlist+1], ids int64 torch tensor on `device`, ascending per list)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    sizes_t = torch.as_tensor(sizes, dtype=torch.int64, device=device)
    perm = torch.randperm(n_total, generator=g, device=device, dtype=torch.int64)
    labels = torch.repeat_interleave(torch.arange(sizes.size, device=device, dtype=torch.int64), sizes_t)
    perm.add_(labels.bitwise_left_shift_(32))  # key = list << 32 | id
    del labels
    key, _ = torch.sort(perm)
    del perm, _
    key.bitwise_and_(0xFFFFFFFF)
    offsets = np.zeros(sizes.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(sizes)
    return offsets, key


def uniform_label_lists(n_total: int, nlist: int, seed: int, device):
    """C2/C4 style: labels = randint(nlist) per id, ids ascending per list."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    labels = torch.randint(0, nlist, (n_total,), generator=g, device=device, dtype=torch.int64)
    key = (labels << 32) | torch.arange(n_total, device=device, dtype=torch.int64)
    sizes = torch.bincount(labels, minlength=nlist).cpu().numpy()
    del labels
    key, _ = torch.sort(key)
    key.bitwise_and_(0xFFFFFFFF)
    offsets = np.zeros(nlist + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(sizes)
    return offsets, key


def nsg_like_graph(n_nodes: int, k: int, seed: int, device):
    """C3: row degree k for 90 % of rows, uniform in [16, k) otherwise; neighbours distinct, != self, -1 padded."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    # distinct neighbours per row: k random keys sorted would be sorted ids; instead draw k+8 candidates,
    # drop duplicates by sorting + masking, then restore a random order.
    cand = torch.randint(0, n_nodes - 1, (n_nodes, k + 8), generator=g, device=device, dtype=torch.int32)
    rows = torch.arange(n_nodes, device=device, dtype=torch.int32).unsqueeze(1)
    cand += (cand >= rows).to(torch.int32)  # skip self
    srt, _ = torch.sort(cand, dim=1)
    dup = torch.zeros_like(srt, dtype=torch.bool)
    dup[:, 1:] = srt[:, 1:] == srt[:, :-1]
    noise = torch.rand(srt.shape, generator=g, device=device)
    noise[dup] = 2.0  # duplicates go last
    order = torch.argsort(noise, dim=1)
    pick = torch.gather(srt, 1, order)[:, :k].contiguous()
    deg = torch.where(torch.rand(n_nodes, generator=g, device=device) < 0.9,
                      torch.full((n_nodes,), k, device=device, dtype=torch.int64),
                      torch.randint(min(16, k - 1), k, (n_nodes,), generator=g, device=device))
    ndup = dup.sum(dim=1)
    deg = torch.minimum(deg, (k + 8 - ndup))
    col = torch.arange(k, device=device).unsqueeze(0)
    pick[col >= deg.unsqueeze(1)] = -1
    return pick, deg
