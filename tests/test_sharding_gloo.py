"""CPU, world_size 2 and 3, gloo: the N>1 path of the ROC codec -- tensor broadcast of the offsets, the
contiguous-unit-range plan, scatter of raw id blocks, per-rank encode, gather-v of the payload tensors, assembly --
must give, on the owner, exactly the payload a single process produces for the whole index. The codec is injected
(the oracle here; the CUDA codec on a GPU box: tests/test_gpu_sharded.py, which needs >= 2 GPUs)."""
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
MAX_UNIT = 1000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def make_index(seed=0, nlist=37):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, 400, size=nlist)
    sizes[0] = 0
    sizes[5] = 0
    sizes[11] = 3000   # three units, straddles a rank boundary
    sizes[12] = 2500
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(sizes)
    ids = np.concatenate([np.sort(rng.choice(1 << 20, size=int(s), replace=False)) for s in sizes]).astype(np.int64)
    return offsets, ids


class OracleCodec:
    """Same interface as sharding.RocCudaCodec, backed by the CPU oracle (test infrastructure)."""

    max_unit = MAX_UNIT

    def __init__(self):
        sys.path.insert(0, str(ROOT))
        import oracle

        self.oracle = oracle

    def encode(self, local_offsets, local_ids):
        from vector_db_id_compression_b200.sharding import unit_table

        ids = local_ids.cpu().numpy().astype(np.uint64)
        _, start, n = unit_table(np.asarray(local_offsets, dtype=np.int64), self.max_unit)
        if np.asarray(local_offsets).size == 1:
            start, n = start[:0], n[:0]
        prec, heads, nwords, lo, hi, words = [], [], [], [], [], []
        for s, m in zip(start.tolist(), n.tolist()):
            if m == 0:
                prec.append(0), heads.append(1 << 31), nwords.append(0), lo.append(0), hi.append(0)
                continue
            seg = ids[s: s + m]
            p = self.oracle.port.precision_rule(int(seg.max()))
            h, w = self.oracle.port.encode(seg, p)
            prec.append(p), heads.append(h), nwords.append(w.size), lo.append(int(seg.min())), hi.append(int(seg.max()))
            words.append(w)
        return dict(precision=np.asarray(prec, np.uint8), heads=np.asarray(heads, np.uint64).view(np.int64),
                    nwords=np.asarray(nwords, np.uint32).view(np.int32), lo=np.asarray(lo, np.uint32).view(np.int32),
                    hi=np.asarray(hi, np.uint32).view(np.int32),
                    words=(np.concatenate(words) if words else np.zeros(0, np.uint32)).view(np.int32))

    def payload(self, blob, device):
        return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in blob.items()}

    def assemble(self, offsets, payload):
        return {k: v.numpy() for k, v in payload.items()}


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vector_db_id_compression_b200 import sharding

    offsets, ids = make_index() if rank == 0 else (None, None)
    phases = []
    whole, local, plan = sharding.encode_sharded(offsets, ids, OracleCodec(), torch.device("cpu"), timer=phases.append)
    assert phases == ["plan", "scatter", "encode", "gather", "assemble"]
    if rank == 0:
        q.put({k: v.tolist() for k, v in whole.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_unit_range_plan_covers_and_balances():
    from vector_db_id_compression_b200.sharding import unit_range_plan, unit_table

    rng = np.random.default_rng(0)
    sizes = np.concatenate([[0, 5, 2500, 0, 1000, 999, 1001, 3000, 70000], rng.integers(0, 400, size=300)])
    off = np.zeros(sizes.size + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    for world in (1, 2, 3, 8):
        plan = unit_range_plan(off, 1000, world)
        assert plan["ucut"][0] == 0 and plan["ucut"][-1] == plan["nunits"] and plan["ecut"][-1] == off[-1]
        # every rank's own list -> unit split reproduces the global one
        want = unit_table(off, 1000)[2]
        got = np.concatenate([unit_table(l.astype(np.int64), 1000)[2] if l.size > 1 else np.zeros(0, np.int64)
                              for l in plan["local_offsets"]])
        assert np.array_equal(want, got)
        loads = np.diff(plan["ecut"])
        assert loads.max() - loads.min() <= 2 * 1000, loads   # within one unit of the ideal cut on either side
        for r in range(world):
            assert int(plan["local_offsets"][r][-1]) == loads[r]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_encode_equals_single_process(world):
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
    want = OracleCodec().encode(offsets, torch.from_numpy(ids))
    for k in want:
        assert np.array_equal(np.asarray(got[k], dtype=want[k].dtype), want[k]), k
