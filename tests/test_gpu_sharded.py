"""GPU: the sharded ROC path. (1) One GPU: the wire form of a blob (idc_roc_blob_export_payload) and its inverse
(idc_roc_blob_assemble), device and host memory; per-"rank" blobs of the contiguous-unit-range plan concatenate to
the blob of the whole index, byte for byte. (2) >= 2 GPUs (skipped otherwise): the same over NCCL -- rank 0 owns the
index, scatter of raw id blocks, per-rank encode, gather-v of device payloads, assembly -- byte-identical to the
1-GPU blob, and the assembled blob decodes to the 1-GPU decode."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def make_index(seed=0, nlist=500, max_unit=4096):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(0, 600, size=nlist)
    sizes[0] = 0
    sizes[7] = 3 * max_unit + 17     # four units: straddles rank boundaries
    sizes[8] = 2 * max_unit
    sizes[400] = 0
    n = int(sizes.sum())
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(sizes)
    perm = rng.permutation(1 << 22)[:n].astype(np.int64)
    ids = np.concatenate([np.sort(perm[int(offsets[l]): int(offsets[l + 1])]) for l in range(nlist)])
    return offsets, ids


def same_blob(a: dict, b: dict):
    for k in ("list_offsets", "unit_offsets", "unit_n", "precision", "heads", "word_offsets", "words"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("mem", ["device", "host"])
def test_payload_roundtrip_and_rankwise_concatenation(mem):
    from vector_db_id_compression_b200.capi import Context
    from vector_db_id_compression_b200.sharding import PAYLOAD_KEYS, unit_range_plan

    ctx = Context(0)
    dev = torch.device("cuda", 0) if mem == "device" else None
    max_unit = 4096
    offsets, ids = make_index(max_unit=max_unit)
    ids_t = torch.as_tensor(ids, device="cuda:0")
    whole = ctx.roc_encode(offsets, ids_t, sorted_ids=True, max_unit=max_unit)
    want = whole.export()
    want_dec, _ = whole.decode()
    pay = whole.export_payload(device=dev)
    again = ctx.roc_assemble(offsets, pay, max_unit=max_unit)
    same_blob(again.export(), want)
    assert again.ans_bytes == whole.ans_bytes and np.array_equal(again.decode()[0], want_dec)
    again.free()
    for world in (2, 3, 8):
        plan = unit_range_plan(offsets.astype(np.int64), max_unit, world)
        parts = []
        for r in range(world):
            e0, e1 = int(plan["ecut"][r]), int(plan["ecut"][r + 1])
            b = ctx.roc_encode(plan["local_offsets"][r], ids_t[e0:e1], sorted_ids=True, max_unit=max_unit)
            assert b.nunits == int(plan["ucut"][r + 1] - plan["ucut"][r])
            parts.append(b.export_payload(device=dev))
            b.free()
        cat = (lambda xs: torch.cat(list(xs))) if dev is not None else (lambda xs: np.concatenate(list(xs)))
        merged = {k: cat(p[k] for p in parts) for k in PAYLOAD_KEYS}
        asm = ctx.roc_assemble(offsets, merged, max_unit=max_unit)
        same_blob(asm.export(), want)
        assert np.array_equal(asm.decode()[0], want_dec)
        asm.free()
    whole.free()
    ctx.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from vector_db_id_compression_b200 import sharding
    from vector_db_id_compression_b200.capi import Context

    max_unit = 4096
    ctx = Context(rank)
    offsets, ids = (None, None)
    if rank == 0:
        offsets, ids = make_index(max_unit=max_unit)
        ids = torch.as_tensor(ids, device=dev)
    whole, local, plan = sharding.encode_sharded(offsets, ids, sharding.RocCudaCodec(ctx, max_unit), dev)
    ok = True
    if rank == 0:
        single = ctx.roc_encode(offsets, ids, sorted_ids=True, max_unit=max_unit)
        a, b = whole.export(), single.export()
        ok = all(np.array_equal(a[k], b[k]) for k in ("list_offsets", "unit_offsets", "unit_n", "precision", "heads",
                                                      "word_offsets", "words"))
        ok = ok and np.array_equal(whole.decode()[0], single.decode()[0]) and whole.ans_bytes == single.ans_bytes
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_over_nccl_is_byte_identical_to_one_gpu(world):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
