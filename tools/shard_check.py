"""C4 of BASELINE.json on N GPUs over NCCL: rank 0 owns an IVF65536 index of 10 M ids; LPT partition, scatter of the
raw id blocks (batched ncclSend / ncclRecv), per-rank ROC encode on the GPU, gather of the blobs; rank 0 checks the
re-assembled tables byte for byte against its own single-GPU encode.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/shard_check.py"""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vector_db_id_compression_b200 import sharding, workloads as W
from vector_db_id_compression_b200.capi import Context

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)

def encode_fn(loc, ids):
    blob = ctx.roc_encode(np.asarray(loc, dtype=np.uint64), ids, sorted_ids=True)
    ex = blob.export()
    blob.free()
    return {k: ex[k] for k in ("unit_offsets", "unit_n", "precision", "heads", "word_offsets", "words")}

offsets = ids = None
if rank == 0:
    offsets, ids_t = W.uniform_label_lists(10_000_000, 65536, 5, dev)
    ids = ids_t
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    res = sharding.encode_sharded(offsets, ids, encode_fn, dev)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    want = encode_fn(offsets, ids)
    ok = all(np.array_equal(np.asarray(res[k]), np.asarray(want[k])) for k in want)
    print(f"shard_check: world {world}, 10 M ids in 65536 lists, scatter + encode + gather {dt*1e3:.0f} ms, "
          f"re-assembled blob {'byte-identical to' if ok else 'DIFFERS FROM'} the single-GPU encode "
          f"({int(np.asarray(want['words']).size)} words)", flush=True)
dist.destroy_process_group()
