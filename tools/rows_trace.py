import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W
dev = torch.device("cuda:0")
data, _ = W.nsg_like_graph(1_000_000, 64, 3, dev)
ctx = Context(0)
for enc in (ctx.ef_encode_rows, ctx.roc_encode_rows):
    b = enc(data); b.free()
    print("---- second call", file=sys.stderr, flush=True)
    b = enc(data); b.free()
    print("====", file=sys.stderr, flush=True)
