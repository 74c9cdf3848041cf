"""CPU, world_size 2, gloo: the N>1 path -- LPT partition, scatter of raw id blocks, per-rank encode,
gather of blobs -- re-assembled result must equal the single-process result byte for byte. The codec is
injected (the oracle here; the CUDA codec on a GPU box: tests/test_gpu_parity.py::test_sharded_matches_single)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_index(seed=0, nlist=37):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, 400, size=nlist)
    sizes[5] = 0
    sizes[11] = 3000
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(sizes)
    ids = np.concatenate([np.sort(rng.choice(1 << 20, size=int(s), replace=False)) for s in sizes]).astype(np.int64)
    return offsets, ids


def oracle_encode_fn(max_unit):
    sys.path.insert(0, str(ROOT))
    import oracle

    def fn(loc, local_ids):
        loc = np.asarray(loc, dtype=np.int64)
        ids = local_ids.cpu().numpy().astype(np.uint64)
        unit_offsets, unit_n, prec, heads, words, woff = [0], [], [], [], [], [0]
        for l in range(loc.size - 1):
            s, e = int(loc[l]), int(loc[l + 1])
            if e == s:
                unit_n.append(0), prec.append(0), heads.append(1 << 31), woff.append(woff[-1])
            for a in range(s, e, max_unit):
                seg = ids[a: min(e, a + max_unit)]
                p = oracle.port.precision_rule(int(seg.max()))
                h, w = oracle.port.encode(seg, p)
                unit_n.append(seg.size), prec.append(p), heads.append(h), words.append(w)
                woff.append(woff[-1] + w.size)
            unit_offsets.append(len(unit_n))
        return dict(unit_offsets=np.asarray(unit_offsets, np.uint64), unit_n=np.asarray(unit_n, np.uint32),
                    precision=np.asarray(prec, np.uint8), heads=np.asarray(heads, np.uint64),
                    word_offsets=np.asarray(woff, np.uint64),
                    words=np.concatenate(words) if words else np.zeros(0, np.uint32))

    return fn


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vector_db_id_compression_b200 import sharding

    offsets, ids = make_index() if rank == 0 else (None, None)
    res = sharding.encode_sharded(offsets, ids, oracle_encode_fn(1000), torch.device("cpu"))
    if rank == 0:
        q.put({k: v.tolist() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_partition_balances_and_covers():
    from vector_db_id_compression_b200.sharding import lpt_partition

    rng = np.random.default_rng(0)
    costs = np.concatenate([rng.integers(1, 100, size=5000), [65536] * 37]).astype(float)
    parts = lpt_partition(costs, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(costs.size))
    loads = np.array([costs[p].sum() for p in parts])
    assert loads.max() / loads.mean() < 1.03


def test_sharded_encode_equals_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    offsets, ids = make_index()
    want = oracle_encode_fn(1000)(offsets, torch.from_numpy(ids))
    for k in want:
        assert np.array_equal(np.asarray(got[k]), want[k]), k
