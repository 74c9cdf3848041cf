"""Sharding of lists / adjacency rows across the GPUs of one box.

Lists are independent coding units (custom_invlists_impl.h:59,76; altid_impl.h:43,58), so the codec needs
no collective: one process per GPU, each encodes / decodes its own lists. NCCL (over NVLink 5 / NVSwitch) is
used only to move data when a single rank owns the index: scatter-v of raw id blocks, gather-v of the
compressed blobs -- device buffers end to end, no pickled objects. All functions work on any torch.distributed
backend (tests run them on gloo / CPU with an injected codec; production passes RocCudaCodec on nccl / CUDA).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


# ---------------------------------------------------------------- ROC: sharding by contiguous unit ranges
# A ROC *unit* (a list of <= max_unit ids, or one max_unit-run of a longer list) is an independent stream. The units
# of an index, in list order, are cut into `world` contiguous ranges of near-equal id count; rank r gets the ids of
# its range -- ONE contiguous slice of the owner's id array, so the scatter is a plain send of slices, no gather
# kernel, no copy -- encodes them, and the owner concatenates the per-rank payloads (idc_roc_blob_export_payload) in
# rank order: that IS the blob of the whole index (idc_roc_blob_assemble). A list longer than max_unit may straddle
# two ranks; both pieces start at a unit boundary, so each rank's own list -> unit split reproduces the global one.
#
# Nothing in the data path is pickled: the plan is a pure function of (offsets, max_unit, world) that every rank
# evaluates after ONE tensor broadcast of the offsets; ids, payload arrays and sizes travel as tensors (NCCL
# send / recv / all_gather over NVLink on CUDA tensors, gloo on CPU tensors in the tests).

def unit_table(offsets: np.ndarray, max_unit: int):
    """(list, start element, n) of every unit in blob order -- the split plan_units_csr (roc_kernels.cu) makes;
    an empty list owns one empty unit."""
    offsets = np.asarray(offsets, dtype=np.int64)
    sizes = np.diff(offsets)
    per_list = np.maximum(1, -(-sizes // max_unit))
    first = np.zeros(sizes.size + 1, dtype=np.int64)
    np.cumsum(per_list, out=first[1:])
    unit_list = np.repeat(np.arange(sizes.size, dtype=np.int64), per_list)
    seg = np.arange(int(first[-1]), dtype=np.int64) - first[unit_list]
    start = offsets[unit_list] + seg * max_unit
    n = np.clip(sizes[unit_list] - seg * max_unit, 0, max_unit)
    return unit_list, start, n


def unit_range_plan(offsets: np.ndarray, max_unit: int, world: int) -> dict:
    """Contiguous unit ranges of near-equal id count. -> ucut[world+1] (unit ranges), ecut[world+1] (element ranges
    into the owner's id array), local_offsets[r] (the CSR rank r encodes, relative to ecut[r])."""
    offsets = np.asarray(offsets, dtype=np.int64)
    unit_list, start, n = unit_table(offsets, max_unit)
    nunits = n.size
    cum = np.cumsum(n)
    total = int(cum[-1]) if nunits else 0
    targets = (np.arange(1, world, dtype=np.float64) * total / world)
    ucut = np.concatenate([[0], np.searchsorted(cum, targets, side="left") + (1 if nunits else 0), [nunits]]).astype(np.int64)
    ucut = np.minimum(np.maximum.accumulate(ucut), nunits)
    ends = start + n
    ecut = np.zeros(world + 1, dtype=np.int64)
    ecut[0] = offsets[0]
    for r in range(world):
        ecut[r + 1] = ends[ucut[r + 1] - 1] if ucut[r + 1] > ucut[r] else ecut[r]
    local = []
    for r in range(world):
        u0, u1 = int(ucut[r]), int(ucut[r + 1])
        if u1 == u0:
            local.append(np.zeros(1, dtype=np.uint64))
            continue
        ul = unit_list[u0:u1]
        grp = np.concatenate([[0], np.nonzero(np.diff(ul))[0] + 1])   # first unit of every (piece of a) list
        sizes = np.add.reduceat(n[u0:u1], grp)
        lo = np.zeros(sizes.size + 1, dtype=np.uint64)
        lo[1:] = np.cumsum(sizes)
        local.append(lo)
    return dict(ucut=ucut, ecut=ecut, local_offsets=local, nunits=nunits, total=total)


def _rank_world(group=None):
    """(rank, world) of the process group; (0, 1) when torch.distributed is not initialised (one GPU, no launcher)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def broadcast_offsets(offsets, max_unit: int, device, src: int = 0, group=None):
    """The owner's CSR offsets (and max_unit) to every rank, as tensors."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world(group)
    if world == 1:
        return np.asarray(offsets).astype(np.int64), int(max_unit)
    head = torch.zeros(2, dtype=torch.int64, device=device)
    if rank == src:
        off_np = np.asarray(offsets).astype(np.int64)
        head = torch.tensor([off_np.size, int(max_unit)], dtype=torch.int64, device=device)
    dist.broadcast(head, src=src, group=group)
    n, mu = int(head[0]), int(head[1])
    t = torch.as_tensor(off_np, device=device) if rank == src else torch.empty(n, dtype=torch.int64, device=device)
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy(), mu


def scatter_id_blocks(plan: dict, ids, device, src: int = 0, group=None):
    """Rank r receives ids[ecut[r] : ecut[r+1]] of the owner's array (the owner's own block is a view of it)."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world(group)
    ecut = plan["ecut"]
    n_mine = int(ecut[rank + 1] - ecut[rank])
    ops = []
    if rank == src:
        ids_t = ids if isinstance(ids, torch.Tensor) else torch.as_tensor(np.asarray(ids, dtype=np.int64))
        ids_t = ids_t.to(device)
        mine = ids_t[int(ecut[rank]): int(ecut[rank + 1])]
        for r in range(world):
            if r != src and ecut[r + 1] > ecut[r]:
                ops.append(dist.P2POp(dist.isend, ids_t[int(ecut[r]): int(ecut[r + 1])], r, group))
    else:
        mine = torch.empty(n_mine, dtype=torch.int64, device=device)
        if n_mine:
            ops.append(dist.P2POp(dist.irecv, mine, src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return mine


PAYLOAD_KEYS = ("precision", "heads", "nwords", "lo", "hi", "words")


def gather_payloads(plan: dict, payload: dict, device, dst: int = 0, group=None) -> Optional[dict]:
    """gather-v of the per-rank payload tensors into pre-sized arrays on `dst`, in rank (= unit) order."""
    import torch
    import torch.distributed as dist

    rank, world = _rank_world(group)
    if world == 1:
        return payload
    ucut = plan["ucut"]
    nw_mine = torch.tensor([int(payload["words"].shape[0])], dtype=torch.int64, device=device)
    nw_all = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(nw_all, nw_mine, group=group)  # blob sizes, as tensors
    nw = np.array([int(t.item()) for t in nw_all], dtype=np.int64)
    wcut = np.concatenate([[0], np.cumsum(nw)])
    ops, out = [], None
    if rank == dst:
        out = {k: torch.empty(int(wcut[-1]) if k == "words" else plan["nunits"], dtype=payload[k].dtype, device=device)
               for k in PAYLOAD_KEYS}
        for r in range(world):
            for k in PAYLOAD_KEYS:
                a, b = (int(wcut[r]), int(wcut[r + 1])) if k == "words" else (int(ucut[r]), int(ucut[r + 1]))
                if b == a:
                    continue
                if r == dst:
                    out[k][a:b].copy_(payload[k])
                else:
                    ops.append(dist.P2POp(dist.irecv, out[k][a:b], r, group))
    else:
        for k in PAYLOAD_KEYS:
            if payload[k].numel():
                ops.append(dist.P2POp(dist.isend, payload[k].contiguous(), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


class RocCudaCodec:
    """The production codec of the sharded path: capi.Context on this rank's GPU."""

    def __init__(self, ctx, max_unit: int = 65536):
        self.ctx, self.max_unit = ctx, max_unit

    def encode(self, local_offsets, local_ids):
        return self.ctx.roc_encode(local_offsets, local_ids, sorted_ids=True, max_unit=self.max_unit)

    def payload(self, blob, device):
        return blob.export_payload(device=device)

    def assemble(self, offsets, payload):
        return self.ctx.roc_assemble(offsets, payload, max_unit=self.max_unit)


def encode_sharded(offsets, ids, codec, device, src: int = 0, group=None, timer=None, resident_ids=None):
    """broadcast offsets -> plan -> scatter id blocks -> per-rank encode -> gather payloads -> assemble on `src`.

    codec: .encode(local_offsets, local_ids) -> blob, .payload(blob, device) -> dict of tensors (PAYLOAD_KEYS),
    .assemble(offsets, payload) -> the whole index's blob. Returns (assembled blob on `src` else None, this rank's
    local blob, plan). timer(name) is called after each phase (bench.py times the phases with it).
    resident_ids: this rank's block is already on the device (skips the scatter)."""
    rank, world = _rank_world(group)
    tick = timer or (lambda name: None)
    goff, max_unit = broadcast_offsets(offsets, getattr(codec, "max_unit", 65536), device, src=src, group=group)
    plan = unit_range_plan(goff, max_unit, world)
    tick("plan")
    mine = resident_ids if resident_ids is not None else scatter_id_blocks(plan, ids, device, src=src, group=group)
    tick("scatter")
    blob = codec.encode(plan["local_offsets"][rank], mine)
    tick("encode")
    payload = codec.payload(blob, device)
    gathered = gather_payloads(plan, payload, device, dst=src, group=group)
    tick("gather")
    whole = codec.assemble(goff, gathered) if rank == src else None
    tick("assemble")
    return whole, blob, plan


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
