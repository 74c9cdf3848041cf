"""Sharding of lists / adjacency rows across the GPUs of one box.

Lists are independent coding units (custom_invlists_impl.h:59,76; altid_impl.h:43,58), so the codec needs
no collective: one process per GPU, each encodes / decodes its own lists. NCCL (over NVLink 5 / NVSwitch) is
used only to move data when a single rank owns the index: scatter-v of raw id blocks, gather-v of the
compressed blobs. All functions work on any torch.distributed backend (tests run them on gloo / CPU with an
injected codec; production passes a capi.Context-based codec on nccl / CUDA).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def roc_cost(n: np.ndarray) -> np.ndarray:
    """Serial steps dominate ROC: cost ~ n (each step is a fixed-latency chain); EF cost ~ n as well."""
    return np.asarray(n, dtype=np.float64)


def lpt_partition(costs: np.ndarray, nparts: int) -> list[np.ndarray]:
    """Longest-processing-time-first greedy partition. Returns, per part, the item indices (ascending)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    loads = np.zeros(nparts)
    parts: list[list[int]] = [[] for _ in range(nparts)]
    # large items individually, the long tail in round-robin blocks (keeps this O(n log n) in numpy terms)
    head = min(order.size, 64 * nparts)
    for i in order[:head]:
        p = int(np.argmin(loads))
        parts[p].append(int(i))
        loads[p] += costs[i]
    tail = order[head:]
    if tail.size:
        # water-filling: every part is topped up to the common target with a run of consecutive tail items
        tc = costs[tail]
        need = np.maximum(costs.sum() / nparts - loads, 0.0)
        need = need / need.sum() * tc.sum() if need.sum() > 0 else np.full(nparts, tc.sum() / nparts)
        cuts = np.searchsorted(np.cumsum(tc), np.cumsum(need)[:-1], side="left")
        for p, seg in enumerate(np.split(tail, cuts)):
            parts[p].extend(int(x) for x in seg)
            loads[p] += costs[seg].sum()
    return [np.sort(np.asarray(p, dtype=np.int64)) for p in parts]


def shard_csr(offsets: np.ndarray, lists: np.ndarray):
    """Sub-CSR of the given lists: (local offsets, gather index ranges)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    sizes = offsets[lists + 1] - offsets[lists]
    loc = np.zeros(lists.size + 1, dtype=np.uint64)
    loc[1:] = np.cumsum(sizes)
    return loc, sizes


def _gather_index(goff: np.ndarray, lists: np.ndarray) -> np.ndarray:
    """Element indices of the given lists, concatenated in the given order (vectorised: no per-list loop)."""
    if lists.size == 0:
        return np.zeros(0, np.int64)
    starts = goff[lists].astype(np.int64)
    sizes = (goff[lists + 1] - goff[lists]).astype(np.int64)
    out_start = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    return np.repeat(starts - out_start, sizes) + np.arange(int(sizes.sum()), dtype=np.int64)


def scatter_lists(offsets, ids, device, src: int = 0, group=None):
    """Rank `src` owns (offsets, ids); every rank gets (its list numbers, local offsets, its ids on `device`).

    Plan = LPT over list lengths, broadcast as an object; raw id blocks move with batched send/recv
    (ncclSend/ncclRecv under nccl) -- one message per destination rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    plan = [None]
    if rank == src:
        offsets = np.asarray(offsets, dtype=np.int64)
        sizes = np.diff(offsets)
        plan = [(lpt_partition(roc_cost(sizes), world), offsets)]
    dist.broadcast_object_list(plan, src=src, group=group)
    parts, goff = plan[0]
    mine = parts[rank]
    loc, sizes = shard_csr(goff, mine)
    n_mine = int(loc[-1])
    recv = torch.empty(n_mine, dtype=torch.int64, device=device)
    ops = []
    if rank == src:
        ids_t = torch.as_tensor(ids, dtype=torch.int64, device=device)
        blocks = []
        for r in range(world):
            lists = parts[r]
            idx = _gather_index(goff, lists)
            blk = ids_t[torch.as_tensor(idx, device=device)] if idx.size else torch.empty(0, dtype=torch.int64, device=device)
            if r == src:
                recv.copy_(blk)
            elif blk.numel():
                blocks.append(blk)
                ops.append(dist.P2POp(dist.isend, blk, r, group))
    elif n_mine:
        ops.append(dist.P2POp(dist.irecv, recv, src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return mine, loc, recv


def gather_blobs(mine: np.ndarray, local: dict, nlist: int, dst: int = 0, group=None) -> Optional[dict]:
    """Re-assemble per-rank ROC exports (capi.RocBlob.export() dicts over the rank's lists) in global list
    order on rank `dst`. Sizes travel with all_gather_object, payload arrays with gather_object (gather-v)."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    payload = dict(lists=np.asarray(mine), **{k: np.asarray(v) for k, v in local.items()})
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst, group=group)
    if rank != dst:
        return None
    # global tables in list order, vectorised: every rank's units / words are scattered to their global positions
    nunits_of = np.zeros(nlist, dtype=np.int64)
    for p in out:
        nunits_of[p["lists"].astype(np.int64)] = np.diff(p["unit_offsets"].astype(np.int64))
    unit_offsets = np.zeros(nlist + 1, np.uint64)
    unit_offsets[1:] = np.cumsum(nunits_of)
    nunits = int(unit_offsets[-1])
    unit_n, precision = np.zeros(nunits, np.uint32), np.zeros(nunits, np.uint8)
    heads, nwords = np.zeros(nunits, np.uint64), np.zeros(nunits, np.int64)
    dest_units = []
    for p in out:
        lists = p["lists"].astype(np.int64)
        uo = p["unit_offsets"].astype(np.int64)
        nu = np.diff(uo)
        du = np.repeat(unit_offsets[lists].astype(np.int64) - uo[:-1], nu) + np.arange(int(nu.sum()), dtype=np.int64)
        dest_units.append(du)
        unit_n[du], precision[du], heads[du] = p["unit_n"][: du.size], p["precision"][: du.size], p["heads"][: du.size]
        nwords[du] = np.diff(p["word_offsets"].astype(np.int64))[: du.size]
    word_offsets = np.zeros(nunits + 1, np.uint64)
    word_offsets[1:] = np.cumsum(nwords)
    words = np.zeros(int(word_offsets[-1]), np.uint32)
    for p, du in zip(out, dest_units):
        wo = p["word_offsets"].astype(np.int64)[: du.size + 1]
        nw = np.diff(wo)
        dw = np.repeat(word_offsets[du].astype(np.int64) - wo[:-1], nw) + np.arange(int(nw.sum()), dtype=np.int64)
        words[dw] = np.asarray(p["words"])[int(wo[0]): int(wo[0]) + dw.size] if dw.size else words[dw]
    return dict(unit_offsets=unit_offsets, unit_n=unit_n, precision=precision, heads=heads,
                word_offsets=word_offsets, words=words)


def encode_sharded(offsets, ids, encode_fn: Callable, device, src: int = 0, group=None) -> Optional[dict]:
    """scatter -> per-rank encode -> gather. encode_fn(local_offsets, local_ids) must return a
    RocBlob.export()-style dict. Returns the global blob tables on rank `src`, None elsewhere."""
    import torch.distributed as dist

    nlist = [None]
    if dist.get_rank(group) == src:
        nlist = [int(np.asarray(offsets).size - 1)]
    dist.broadcast_object_list(nlist, src=src, group=group)
    mine, loc, local_ids = scatter_lists(offsets, ids, device, src=src, group=group)
    local = encode_fn(loc, local_ids)
    return gather_blobs(mine, local, nlist[0], dst=src, group=group)


# ---------------------------------------------------------------- wavelet tree: sharding by id range
# The wavelet-tree index is ONE structure over S[id] = list_no (custom_invlists_impl.cpp:346-397), so it cannot be
# sharded by lists. It shards by ID RANGE: rank r indexes the ids [lo[r], lo[r+1]) (shifted to start at 0), i.e. the
# slice S[lo[r] : lo[r+1]]. Ids are ascending inside a list, so list c's ids are the concatenation over ranks of its
# local ids, and a (world x nlist) table of per-rank list counts routes get_single_id(c, k) to the one rank that
# holds it. No collective inside the codec; NCCL moves the raw id blocks (scatter) and the answers (all-reduce /
# gather) only.

def wt_id_range_plan(offsets, ids, world: int, align: int = 512) -> dict:
    """Split an index whose lists partition [0, n) into `world` id ranges of `chunk` ids (a multiple of `align`).
    ids: numpy array or torch tensor (the plan is computed where the ids live: on the GPU for a device tensor).
    -> lo[world+1], counts[world, nlist] (ids of list c held by rank r; numpy), order (torch: element indices
    grouped by rank, list order and id order preserved)."""
    import torch

    offsets = np.asarray(offsets, dtype=np.int64)
    ids_t = torch.as_tensor(ids).to(torch.int64)
    nlist, n = offsets.size - 1, int(offsets[-1] - offsets[0])
    chunk = max(-(-max(n, 1) // world), 1)
    chunk = -(-chunk // align) * align
    lo = np.minimum(np.arange(world + 1, dtype=np.int64) * chunk, n)
    shard = torch.div(ids_t, chunk, rounding_mode="floor")
    sizes = torch.as_tensor(np.diff(offsets), device=ids_t.device)
    lists = torch.repeat_interleave(torch.arange(nlist, device=ids_t.device, dtype=torch.int64), sizes)
    counts = torch.bincount(shard * nlist + lists, minlength=world * nlist).reshape(world, nlist).cpu().numpy()
    order = torch.argsort(shard, stable=True)
    return dict(lo=lo, counts=counts, order=order, chunk=chunk)


class WtShardedIndex:
    """Wavelet-tree index sharded by id range over the ranks of `group`.

    encode_fn(local_offsets[nlist+1] u64, local_ids int64 tensor on `device`) -> an object with the capi.WtBlob
    interface: .select(list_nos, offsets_in_list) -> int64 array, .decode() -> (ids, offsets). Rank `src` owns
    (offsets, ids); the other ranks pass None."""

    def __init__(self, offsets, ids, encode_fn: Callable, device, src: int = 0, group=None):
        import torch
        import torch.distributed as dist

        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        meta = [None]
        plan = None
        if self.rank == src:
            plan = wt_id_range_plan(offsets, ids, self.world)
            meta = [(np.asarray(offsets, dtype=np.int64) - int(offsets[0]), plan["lo"], plan["counts"])]
        dist.broadcast_object_list(meta, src=src, group=group)
        self.offsets, self.lo, self.counts = meta[0]
        self.nlist = self.offsets.size - 1
        # prefix[r, c] = ids of list c held by the ranks before r; prefix[world, c] = size of list c
        self.prefix = np.zeros((self.world + 1, self.nlist), dtype=np.int64)
        np.cumsum(self.counts, axis=0, out=self.prefix[1:])
        n_mine = int(self.counts[self.rank].sum())
        recv = torch.empty(n_mine, dtype=torch.int64, device=device)
        ops, keep = [], []
        if self.rank == src:
            ids_t = torch.as_tensor(ids).to(torch.int64)[plan["order"]].to(device)
            ends = np.cumsum(self.counts.sum(axis=1))
            for r in range(self.world):
                blk = ids_t[int(ends[r]) - int(self.counts[r].sum()): int(ends[r])] - int(self.lo[r])
                if r == src:
                    recv.copy_(blk)
                elif blk.numel():
                    keep.append(blk)
                    ops.append(dist.P2POp(dist.isend, blk, r, group))
        elif n_mine:
            ops.append(dist.P2POp(dist.irecv, recv, src, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.local_offsets = np.zeros(self.nlist + 1, dtype=np.uint64)
        self.local_offsets[1:] = np.cumsum(self.counts[self.rank])
        self.local = encode_fn(self.local_offsets, recv) if n_mine else None

    def select(self, list_nos, offsets_in_list) -> np.ndarray:
        """get_single_id for the same (list, offset) pairs on every rank; -1 outside the index."""
        import torch
        import torch.distributed as dist

        ln = np.asarray(list_nos, dtype=np.int64)
        of = np.asarray(offsets_in_list, dtype=np.int64)
        ok = (ln >= 0) & (ln < self.nlist) & (of >= 0)
        lc = np.where(ok, ln, 0)
        ok &= of < self.prefix[self.world, lc]
        owner = (self.prefix[1:, lc] <= of[None, :]).sum(axis=0)  # ranks whose share of the list ends at or before k
        mine = ok & (owner == self.rank)
        out = np.full(ln.size, -1, dtype=np.int64)
        if mine.any():
            loc = self.local.select(lc[mine], of[mine] - self.prefix[self.rank, lc[mine]])
            out[mine] = np.asarray(loc, dtype=np.int64) + int(self.lo[self.rank])
        t = torch.as_tensor(out, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t.cpu().numpy()

    def decode_all(self, dst: int = 0) -> Optional[np.ndarray]:
        """get_ids of every list, in the global CSR order, on rank `dst` (None elsewhere)."""
        import torch.distributed as dist

        loc = np.zeros(0, dtype=np.int64)
        if self.local is not None:
            loc = np.asarray(self.local.decode()[0], dtype=np.int64) + int(self.lo[self.rank])
        out = [None] * self.world if self.rank == dst else None
        dist.gather_object(loc, out, dst=dst, group=self.group)
        if self.rank != dst:
            return None
        res = np.empty(int(self.offsets[-1]), dtype=np.int64)
        for r, blk in enumerate(out):
            sizes = self.counts[r]
            local_start = np.concatenate([[0], np.cumsum(sizes)[:-1]])
            dest = np.repeat(self.offsets[:-1] + self.prefix[r] - local_start, sizes) + np.arange(blk.size, dtype=np.int64)
            res[dest] = blk
        return res
