"""Wavelet-tree index sharded by id range on N GPUs over NCCL (sharding.WtShardedIndex): rank 0 owns 10^8 ids in
65 536 lists; id-range plan on the device, scatter of the raw id blocks, per-rank idc_wt_encode, routed
get_single_id (all-reduce) and re-assembled get_ids (gather) checked against the input.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/wt_shard_check.py [n_ids]
NOT YET RUN on a GPU box (round 1 ran out of GPU minutes); the same class is covered on gloo / CPU with the oracle as
the per-rank index (tests/test_sharding_wt_gloo.py)."""
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
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
offsets = ids = None
if rank == 0:
    offsets, ids = W.random_partition_lists(n, W.zipf_sizes(n, 65536, 0.0), 5, dev)
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
idx = sharding.WtShardedIndex(offsets, ids, lambda off, loc_ids: ctx.wt_encode(off, loc_ids), dev)
dist.barrier(); torch.cuda.synchronize(); t_build = time.perf_counter() - t0
rng = np.random.default_rng(1)
ql = rng.integers(0, 65536, size=1_000_000)
qo = (rng.random(ql.size) * idx.prefix[world, ql]).astype(np.int64)
t0 = time.perf_counter(); got = idx.select(ql, qo); t_sel = time.perf_counter() - t0
t0 = time.perf_counter(); dec = idx.decode_all(); t_dec = time.perf_counter() - t0
if rank == 0:
    ids_h = ids.cpu().numpy()
    ok_sel = np.array_equal(got, ids_h[offsets[ql].astype(np.int64) + qo])
    ok_dec = np.array_equal(dec, ids_h)
    print(f"wt_shard_check: world {world}, {n} ids in 65536 lists by id range: plan + scatter + build {t_build*1e3:.0f} ms, "
          f"1 M routed selects {t_sel*1e3:.0f} ms ({'ok' if ok_sel else 'MISMATCH'}), gather of all ids {t_dec*1e3:.0f} ms "
          f"({'ok' if ok_dec else 'MISMATCH'})", flush=True)
dist.destroy_process_group()
