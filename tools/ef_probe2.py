"""EF encode / decode on equal-length lists of different counts (field width l follows the list length): python tools/ef_probe2.py"""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W
dev = torch.device("cuda:0")
n = 1_000_000_000
ctx = Context(0); ctx.set_timing(True)
for nlist in (64, 1024, 16384, 65536, 262144, 976563):
    sizes = W.zipf_sizes(n, nlist, 0.0)
    off, ids = W.random_partition_lists(n, sizes, 1234, dev)
    for it in range(3):
        eb = ctx.ef_encode(off, ids, sorted_ids=True)
        be = dict(ctx.last_kernel_breakdown())
        out, _ = eb.decode(device=dev)
        bd = dict(ctx.last_kernel_breakdown())
        ok = bool(torch.equal(out, ids)); l = int(eb.export()["l"][0]) if it == 2 else -1
        eb.free(); del out
    print(json.dumps({"nlist": nlist, "per_list": int(sizes[0]), "l": l, "enc_ms": round(be["k_ef_encode"], 3), "dec_ms": round(bd["k_ef_decode"], 3), "ok": ok}), flush=True)
    del off, ids
