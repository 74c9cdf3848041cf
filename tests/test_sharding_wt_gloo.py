"""CPU, world_size 2 and 3, gloo: the wavelet-tree index sharded by ID RANGE (sharding.WtShardedIndex) -- plan,
scatter of the raw id blocks, per-rank index, routed get_single_id, re-assembled get_ids -- equals the single-process
answers. The per-rank index is injected (the oracle here; capi.Context.wt_encode on a GPU box)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_sharding_gloo import _free_port

ROOT = Path(__file__).resolve().parents[1]


def make_index(seed=3, nlist=23, n=20_000):
    rng = np.random.default_rng(seed)
    lab = rng.integers(0, nlist, size=n)
    lab[lab == 4] = 5          # an empty list
    lab[:3000] = 7             # a list that lives almost entirely in the first id range
    order = np.argsort(lab, kind="stable")
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(np.bincount(lab, minlength=nlist))
    return offsets, order.astype(np.int64)


class OracleWt:
    """capi.WtBlob stand-in on top of oracle/wt_oracle.c"""

    def __init__(self, local_offsets, local_ids):
        sys.path.insert(0, str(ROOT))
        import oracle

        self.o = oracle
        self.off = np.asarray(local_offsets, dtype=np.uint64)
        self.ids = local_ids.cpu().numpy().astype(np.int64)
        self.S = oracle.wt.sequence(self.off, self.ids)  # also checks that the slice is a partition of [0, n_local)
        self.wt = oracle.wt.build(self.off.size - 1, self.S)

    def select(self, list_nos, offs):
        return np.array([self.o.wt.select(self.wt, int(c), int(k)) for c, k in zip(list_nos, offs)], dtype=np.int64)

    def decode(self):
        return self.ids.copy(), self.off


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vector_db_id_compression_b200 import sharding

    offsets, ids = make_index() if rank == 0 else (None, None)
    if rank == 0 and world == 3:
        ids = torch.from_numpy(ids)  # the owner may hold the ids as a (device) tensor
    idx = sharding.WtShardedIndex(offsets, ids, OracleWt, torch.device("cpu"))
    rng = np.random.default_rng(1)  # the same queries on every rank
    ql = rng.integers(-1, 24, size=400)
    qo = rng.integers(0, 1200, size=400)
    got = idx.select(ql, qo)
    dec = idx.decode_all()
    if rank == 0:
        q.put(dict(select=got.tolist(), ql=ql.tolist(), qo=qo.tolist(), decode=dec.tolist(),
                   counts=idx.counts.tolist(), lo=idx.lo.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_id_range_plan_covers_and_preserves_order():
    from vector_db_id_compression_b200.sharding import wt_id_range_plan

    offsets, ids = make_index()
    for world in (1, 2, 3, 8):
        p = wt_id_range_plan(offsets, ids, world)
        assert p["lo"][0] == 0 and p["lo"][-1] == ids.size and p["chunk"] % 512 == 0
        assert np.array_equal(p["counts"].sum(axis=0), np.diff(offsets.astype(np.int64)))
        start = 0
        for r in range(world):
            blk = ids[p["order"].numpy()[start: start + int(p["counts"][r].sum())]]
            start += blk.size
            assert blk.size == p["lo"][r + 1] - p["lo"][r]                 # an id range holds exactly its ids
            assert np.array_equal(np.sort(blk), np.arange(p["lo"][r], p["lo"][r + 1]))
            loc = np.concatenate([[0], np.cumsum(p["counts"][r])])
            for c in range(offsets.size - 1):                               # ascending inside every local list
                assert np.all(np.diff(blk[loc[c]: loc[c + 1]]) > 0)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_wavelet_index_equals_single_process(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    offsets, ids = make_index()
    assert np.array_equal(np.asarray(got["decode"]), ids)
    sizes = np.diff(offsets.astype(np.int64))
    for c, k, g in zip(got["ql"], got["qo"], got["select"]):
        want = int(ids[int(offsets[c]) + k]) if 0 <= c < sizes.size and k < sizes[c] else -1
        assert g == want, (c, k, g, want)
    assert np.asarray(got["counts"]).shape == (world, sizes.size)
