"""C4 of BASELINE.json on N GPUs over NCCL: rank 0 owns an IVF65536 index of 10 M ids; contiguous-unit-range plan,
scatter of the raw id blocks (batched ncclSend / ncclRecv of slices), per-rank ROC encode on the GPU, gather-v of the
device payloads, assembly; rank 0 checks the assembled blob byte for byte against its own single-GPU encode.
bench.py --gpus N reports the same path with timings ("sharded"); this is the stand-alone check.
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
codec = sharding.RocCudaCodec(ctx)
offsets = ids = None
if rank == 0:
    offsets, ids = W.uniform_label_lists(10_000_000, 65536, 5, dev)
for rep in range(2):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    whole, local_blob, plan = sharding.encode_sharded(offsets, ids, codec, dev)
    dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
if rank == 0:
    single = ctx.roc_encode(offsets, ids, sorted_ids=True)
    a, b = whole.export_payload(device=dev), single.export_payload(device=dev)
    ok = all(torch.equal(a[k], b[k]) for k in a)
    print(f"shard_check: world {world}, 10 M ids in 65536 lists, scatter + encode + gather {dt*1e3:.1f} ms, "
          f"assembled blob {'byte-identical to' if ok else 'DIFFERS FROM'} the single-GPU encode "
          f"({single.total_words} words)", flush=True)
dist.destroy_process_group()
