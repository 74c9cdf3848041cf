"""C3 of BASELINE.json at full size: NSG-like adjacency, 1 M rows x K = 64, int32 -- Elias-Fano and ROC row encode,
decode of all rows, and random access to 10 M rows drawn with replacement (device-resident), with a round-trip check."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W

N, K = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000, 64
dev = torch.device("cuda:0")
data, _ = W.nsg_like_graph(N, K, 3, dev)
deg = (data >= 0).sum(1)
edges = int(deg.sum())
ctx = Context(0); ctx.set_timing(True)
sel = torch.randint(0, N, (10_000_000,), generator=torch.Generator(device=dev).manual_seed(4), device=dev, dtype=torch.int32)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); ctx.synchronize(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return r, best
srt = torch.sort(torch.where(data >= 0, data, torch.full_like(data, 2**31 - 1)), dim=1)[0]
for name, enc in (("EF", ctx.ef_encode_rows), ("ROC", ctx.roc_encode_rows)):
    blob, te = timed(lambda: enc(data), reps=2)
    (nb, cnt), td = timed(lambda: blob.decode_rows(device=dev))
    (nb2, cnt2), tr = timed(lambda: blob.decode_rows(sel, device=dev), reps=2)
    got = torch.sort(torch.where(nb >= 0, nb, torch.full_like(nb, 2**31 - 1)), dim=1)[0]
    ok = bool(torch.equal(got, srt)) and bool(torch.equal(cnt.long(), deg)) and bool(torch.equal(nb2, nb[sel.long()]))
    size = blob.bits_total / 8 if name == "EF" else blob.ans_bytes
    print(f"{name}: {N} rows, {edges} edges, {8*size/edges:.2f} bits/edge | encode {te*1e3:.1f} ms ({edges/te/1e9:.2f} G edges/s) | "
          f"decode all rows {td*1e3:.1f} ms ({edges/td/1e9:.2f} G edges/s) | random access 10 M rows {tr*1e3:.1f} ms "
          f"({1e7/tr/1e6:.1f} M rows/s) | round trip {'ok' if ok else 'MISMATCH'}", flush=True)
    blob.free()
