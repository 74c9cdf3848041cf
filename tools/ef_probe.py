"""Elias-Fano kernels alone on the C5 workload (1e9 ids, 65 536 Zipf-length lists): event-timed encode / decode,
round trip; `python tools/ef_probe.py [n_ids] [zipf_s]`."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vector_db_id_compression_b200.capi import Context
from vector_db_id_compression_b200 import workloads as W

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
s = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda:0")
sizes = W.zipf_sizes(n, 65536, s)
off, ids = W.random_partition_lists(n, sizes, 1234, dev)
ctx = Context(0); ctx.set_timing(True)
peak = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))["hbm_gbs"] if (Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").exists() else 6531.9
res = {}
for it in range(6):
    t0 = time.perf_counter()
    eb = ctx.ef_encode(off, ids, sorted_ids=True)
    wall_e = time.perf_counter() - t0
    be = dict(ctx.last_kernel_breakdown())
    t0 = time.perf_counter()
    out, _ = eb.decode(device=dev)
    wall_d = time.perf_counter() - t0
    bd = dict(ctx.last_kernel_breakdown())
    comp = eb.bits_total / 8
    ok = bool(torch.equal(out, ids))
    eb.free(); del out
    res = {"encode_ms": be["k_ef_encode"], "encode_breakdown": be, "decode_ms": bd["k_ef_decode"], "ok": ok,
           "encode_frac": (8 * n + comp) / (be["k_ef_encode"] * 1e-3) / 1e9 / peak, "decode_frac": (8 * n + comp) / (bd["k_ef_decode"] * 1e-3) / 1e9 / peak,
           "encode_wall_ms": 1e3 * wall_e, "decode_wall_ms": 1e3 * wall_d}
print(json.dumps(res))
