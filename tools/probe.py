"""Quick device-side probe: time encode/decode of a synthetic list set with per-kernel breakdown."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=float, default=1e8)
ap.add_argument("--nlist", type=int, default=65536)
ap.add_argument("--zipf", type=float, default=0.0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--ef", type=int, default=1)
ap.add_argument("--roc", type=int, default=1)
a = ap.parse_args()
N = int(a.n)
dev = torch.device("cuda:0")
t0 = time.time()
sizes = W.zipf_sizes(N, a.nlist, a.zipf)
offsets, ids = W.random_partition_lists(N, sizes, 1, dev)
torch.cuda.synchronize()
print(f"gen {time.time()-t0:.2f}s N={N} nlist={a.nlist} zipf={a.zipf} max_list={sizes.max()}", flush=True)
ctx = Context(0)
ctx.set_timing(True)
def timed(fn):
    torch.cuda.synchronize(); t = time.time(); r = fn(); ctx.synchronize(); return r, time.time() - t
if a.roc:
    for rep in range(a.reps):
        blob, te = timed(lambda: ctx.roc_encode(offsets, ids, sorted_ids=True))
        be = ctx.last_kernel_breakdown()
        (out, _), td = timed(lambda: blob.decode(device=dev))
        bd = ctx.last_kernel_breakdown()
        ok = bool((torch.sort(out[: int(offsets[1])])[0] == ids[: int(offsets[1])]).all())
        print(json.dumps(dict(kind="roc", rep=rep, enc_s=te, dec_s=td, enc_Gids=N/te/1e9, dec_Gids=N/td/1e9,
              bytes_per_id=blob.ans_bytes/N, nunits=blob.nunits, enc_kernels=be, dec_kernels=bd, first_list_ok=ok)), flush=True)
        del out
        if rep < a.reps - 1: blob.free()
    blob.free()
if a.ef:
    for rep in range(a.reps):
        eb, te = timed(lambda: ctx.ef_encode(offsets, ids, sorted_ids=True))
        be = ctx.last_kernel_breakdown()
        (out, _), td = timed(lambda: eb.decode(device=dev))
        bd = ctx.last_kernel_breakdown()
        ok = bool((out == ids).all())
        print(json.dumps(dict(kind="ef", rep=rep, enc_s=te, dec_s=td, enc_Gids=N/te/1e9, dec_Gids=N/td/1e9,
              bytes_per_id=eb.bits_total/8/N, enc_kernels=be, dec_kernels=bd, exact=ok)), flush=True)
        del out
        eb.free()
