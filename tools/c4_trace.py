"""C4 (IVF65536, 10^7 ids): wall time of an encode + decode pair and where the host spends it (IDC_TRACE_HOST=1)."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W
dev = torch.device("cuda:0")
off, ids = W.uniform_label_lists(10_000_000, 65536, 5, dev)
ctx = Context(0)
out = torch.empty_like(ids)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    blob = ctx.roc_encode(off, ids, sorted_ids=True)
    t1 = time.perf_counter()
    d, _ = blob.decode(device="cuda", out=out)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("rep %d: encode %.3f ms decode %.3f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3), file=sys.stderr)
    blob.free()
