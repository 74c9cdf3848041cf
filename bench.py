#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native id codec.

Metric (BASELINE.json): ids/s through ROC encode + decode, bit-exact, on the C5 workload
("1 B synthetic IDs in 65 536 Zipf-length lists"), reported next to the achieved fraction of the
HBM roofline and the reference's CPU codec timed on the same box.

One "step" = one pass of the hot path over the whole workload: ROC-encode every list, then ROC-decode
every list. `value` = ids / (encode + decode time) with the ids resident in HBM; `e2e` = the same
through the C ABI with HOST buffers (pinned), H2D of the ids and D2H of the decoded ids inside the
timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...      (weak scaling: one workload per rank)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-ids", type=float, default=1e9, help="ids per GPU (C5: 1e9)")
    ap.add_argument("--nlist", type=int, default=65536)
    ap.add_argument("--zipf-s", type=float, default=1.0, help="list-length exponent; 0 = equal-length control")
    ap.add_argument("--max-unit", type=int, default=65536)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per reference step / baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ef", action="store_true")
    ap.add_argument("--no-wt", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    return ap.parse_args()


# ------------------------------------------------------------------ helpers

def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic(kernel: str, n_ids: int, zipf_s: float):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_traffic.py), summed over the kernel's size-class launches.
    Only reported when the capture was made on this very workload."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        t = json.load(open(p))
        if int(t["n_ids"]) == int(n_ids) and abs(float(t["zipf_s"]) - float(zipf_s)) < 1e-9 and kernel in t["kernels"]:
            return float(t["kernels"][kernel]["dram_bytes"]), t.get("source")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(args, device, seed):
    from vector_db_id_compression_b200 import workloads as W

    n = int(args.n_ids)
    sizes = W.zipf_sizes(n, args.nlist, args.zipf_s)
    offsets, ids = W.random_partition_lists(n, sizes, seed, device)
    return sizes, offsets, ids


def unit_table(offsets: np.ndarray, max_unit: int):
    """(start, n) of every ROC unit, in blob order (mirrors the library's list -> unit split)."""
    starts, ns = [], []
    for l in range(offsets.size - 1):
        s, e = int(offsets[l]), int(offsets[l + 1])
        if e == s:
            starts.append(s)
            ns.append(0)
            continue
        for a in range(s, e, max_unit):
            starts.append(a)
            ns.append(min(max_unit, e - a))
    return np.asarray(starts, dtype=np.int64), np.asarray(ns, dtype=np.int64)


def pick_sample_units(ns: np.ndarray, budget_ids: int, rng, always=()):
    order = rng.permutation(ns.size)
    chosen = list(always)
    have = int(ns[list(always)].sum()) if len(always) else 0
    seen = set(always)
    for u in order:
        if have >= budget_ids:
            break
        if u in seen or ns[u] == 0:
            continue
        chosen.append(int(u))
        have += int(ns[u])
    return np.asarray(chosen, dtype=np.int64)


def cpu_codec():
    import oracle

    if oracle.ref is not None:
        return oracle.ref, "reference"
    return oracle.port, "port"


def cpu_roundtrip(codec, sample_ids: np.ndarray, sample_off: np.ndarray, prec: np.ndarray, threads: int):
    """Reference CPU path on the sample: plugin-style encode loop + get_ids decode. -> (t_enc, t_dec, blobs)."""
    t0 = time.perf_counter()
    heads, nwords, woff, words = codec.encode_lists(sample_off, sample_ids, prec, nthreads=threads)
    t1 = time.perf_counter()
    dec = codec.decode_lists(sample_off, prec, woff, nwords, heads, words, nthreads=threads)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, (heads, nwords, woff, words, dec)


def precision_rule_np(max_ids: np.ndarray) -> np.ndarray:
    # ceil(log2(m)) == bit_length(m - 1) for m >= 1 (custom_invlists_impl.cpp:163-164)
    m = np.maximum(max_ids.astype(np.int64) - 1, 0)
    out = np.zeros(m.size, dtype=np.uint8)
    nz = m > 0
    out[nz] = np.floor(np.log2(m[nz].astype(np.float64))).astype(np.uint8) + 1
    # exact fix-up for float rounding near powers of two
    for i in np.nonzero(nz)[0]:
        out[i] = int(m[i]).bit_length()
    return out


# ------------------------------------------------------------------ reference arm

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    codec, kind = cpu_codec()
    threads = os.cpu_count() or 1
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    sizes, offsets, ids = make_workload(args, dev, args.seed)
    starts, ns = unit_table(offsets, args.max_unit)
    rng = np.random.default_rng(args.seed)

    def gather(units):
        soff = np.zeros(units.size + 1, dtype=np.uint64)
        soff[1:] = np.cumsum(ns[units])
        idx = torch.cat([torch.arange(int(starts[u]), int(starts[u] + ns[u]), device=dev) for u in units])
        sid = ids[idx].cpu().numpy().astype(np.uint64)
        prec = precision_rule_np(np.array([sid[int(soff[i + 1]) - 1] for i in range(units.size)]))
        return sid, soff, prec

    # pilot to size the per-step sample
    pilot = pick_sample_units(ns, max(200_000, 4 * int(ns.max())), rng)
    sid, soff, prec = gather(pilot)
    te, td, _ = cpu_roundtrip(codec, sid, soff, prec, threads)
    rate = sid.size / (te + td)
    budget = int(max(sid.size, rate * args.cpu_seconds))
    units = pick_sample_units(ns, budget, rng)
    sid, soff, prec = gather(units)
    del ids
    times = []
    for step in range(args.warmup + args.steps):
        te, td, _ = cpu_roundtrip(codec, sid, soff, prec, threads)
        if step >= args.warmup:
            times.append((te, td))
    t = float(np.sum(times))
    value = sid.size * args.steps / t
    sample = (f"{units.size} of {ns.size} ROC units chosen uniformly at random ({sid.size} of {int(args.n_ids)} ids) "
              f"per step; encode = plugin loop custom_invlists_impl.cpp:147-194, decode = get_ids :210-219")
    line = {
        "impl": "reference", "metric": "ROC encode+decode ids/s (bit-exact round trip)", "value": value,
        "unit": "ids/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, sizes),
        "cpu_baseline": {"value": value, "unit": "ids/s", "cores": threads, "kind": kind, "sample": sample,
                         "encode_ids_per_s": sid.size * args.steps / float(np.sum([x[0] for x in times])),
                         "decode_ids_per_s": sid.size * args.steps / float(np.sum([x[1] for x in times]))},
        "e2e": {"value": value, "unit": "ids/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sizes):
    return {
        "workload": f"C5: {int(args.n_ids)} synthetic ids per GPU in {args.nlist} lists, lengths ~ k^-{args.zipf_s} "
                    f"(longest {int(sizes.max())}), ids = random partition of [0, N), ascending per list; "
                    f"ROC unit = <= {args.max_unit} consecutive ids of a list",
        "n_ids_per_gpu": int(args.n_ids), "nlist": args.nlist, "zipf_s": args.zipf_s, "max_unit": args.max_unit,
        "l2_policy": "inputs (8 B/id) and outputs exceed the 126 MB L2 by >50x; no explicit flush",
        "parallelism": "lists sharded across GPUs, no data-path collective",
    }


# ------------------------------------------------------------------ B200 arm

def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from vector_db_id_compression_b200.capi import Context

    sizes, offsets, ids = make_workload(args, dev, args.seed + 7919 * rank)
    n_ids = int(ids.numel())
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_timing(True)

    # ---------------- device-resident timed region (value)
    kern_ms = {}

    def one_step(collect: bool):
        blob = ctx.roc_encode(offsets, ids, sorted_ids=True, max_unit=args.max_unit)
        if collect:
            for k, v in ctx.last_kernel_breakdown():
                kern_ms.setdefault(k, []).append(v)
        out, _ = blob.decode(device=dev)
        if collect:
            for k, v in ctx.last_kernel_breakdown():
                kern_ms.setdefault(k, []).append(v)
        return blob, out

    blob = out = None
    for _ in range(args.warmup):
        if blob is not None:
            blob.free()
        blob, out = one_step(False)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        if blob is not None:
            blob.free()
        del out
        blob, out = one_step(True)
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    t_dev = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_max = float(t_dev.item())
    value = n_ids * world * args.steps / (ms_max * 1e-3)

    # ---------------- parity: sampled units (+ the longest) against the CPU oracle; also the cpu_baseline
    info = dict(nunits=blob.nunits, ans_bytes=blob.ans_bytes, total_words=blob.total_words)
    parity = cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        parity, cpu_base = parity_and_cpu_baseline(args, blob, out, offsets, ids, dev)
    elif rank == 0:
        parity = {"checked": False}

    # ---------------- roofline of the dominant kernel (algorithmic bytes / event time)
    peak, peak_src = measured_peak_gbs()
    avg = {k: float(np.mean(v)) for k, v in kern_ms.items()}
    enc_bytes = 8.0 * n_ids + blob.ans_bytes          # read int64 ids, write streams
    dec_bytes = blob.ans_bytes + 8.0 * n_ids          # read streams, write int64 ids
    dom = max(("k_roc_encode", "k_roc_decode"), key=lambda k: avg.get(k, 0.0))
    dom_bytes = enc_bytes if dom == "k_roc_encode" else dec_bytes
    achieved = dom_bytes / (avg[dom] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom, n_ids, args.zipf_s)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": avg[dom],
                "other": {k: {"ms": avg[k]} for k in avg if k != dom}}
    for k, b in (("k_roc_encode", enc_bytes), ("k_roc_decode", dec_bytes)):
        if k in avg and k != dom:
            roofline["other"][k].update(achieved=b / (avg[k] * 1e-3) / 1e9, frac=b / (avg[k] * 1e-3) / 1e9 / peak)
    blob.free()
    del out

    # ---------------- Elias-Fano (the HBM-bound codec) on the same lists
    ef = None
    if not args.no_ef:
        ef = ef_section(args, ctx, offsets, ids, dev, peak)

    # ---------------- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # every rank pins 16 GB of host memory for this leg; if the box cannot give that to all ranks the leg is
        # reported as failed (on every rank alike, so that the collectives inside stay matched) instead of taking
        # the device-resident numbers down with it
        ok_local = 1
        host_bufs = None
        try:
            host_bufs = e2e_buffers(ids)
        except Exception as ex:  # noqa: BLE001
            ok_local = 0
            print(f"bench.py: rank {rank}: pinned host buffers for the e2e leg failed: {ex}", file=sys.stderr)
        flag = torch.tensor([ok_local], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            e2e = e2e_section(args, ctx, offsets, ids, world, dev, barrier, host_bufs)
        else:
            e2e = {"value": None, "unit": "ids/s", "error": "pinned host memory for the e2e leg not available on every rank"}

    # ---------------- wavelet tree (the third id index of the plugin surface) on the same lists (last: a failure here cannot touch the legs above)
    wt = None
    if not args.no_wt:
        try:
            wt = wt_section(args, ctx, offsets, ids, sizes, dev, peak)
        except Exception as ex:  # noqa: BLE001 -- never takes the headline numbers down with it
            wt = {"error": str(ex)}

    if rank == 0:
        line = {
            "metric": "ROC encode+decode ids/s (bit-exact round trip)", "value": value, "unit": "ids/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, sizes),
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity,
            "roc": {"encode_ids_per_s": n_ids / (sum(avg.get(k, 0) for k in ("k_unit_meta", "k_enc_records", "k_roc_encode", "k_roc_compact")) * 1e-3),
                    "decode_ids_per_s": n_ids / (sum(avg.get(k, 0) for k in ("memset_ws", "k_roc_decode")) * 1e-3),
                    "bits_per_id": 8.0 * info["ans_bytes"] / n_ids, "units": info["nunits"],
                    "wall_ms_per_step": 1e3 * t_wall / args.steps},
            "ef": ef, "wt": wt,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_and_cpu_baseline(args, blob, out, offsets, ids, dev):
    """Bit-exact check of a sample of units against the CPU codec, which is timed on the way (cpu_baseline)."""
    import torch

    codec, kind = cpu_codec()
    threads = os.cpu_count() or 1
    ex = blob.export()
    starts, ns = unit_table(offsets, args.max_unit)
    assert ns.size == blob.nunits
    rng = np.random.default_rng(99)
    longest = np.argsort(-ns, kind="stable")[:32]

    def gather(units):
        soff = np.zeros(units.size + 1, dtype=np.uint64)
        soff[1:] = np.cumsum(ns[units])
        idx = torch.cat([torch.arange(int(starts[u]), int(starts[u] + ns[u]), device=dev) for u in units])
        sid = ids[idx].cpu().numpy().astype(np.uint64)
        dec = out[idx].cpu().numpy().astype(np.uint64)
        return sid, soff, dec

    # pilot -> budget
    pilot = pick_sample_units(ns, max(200_000, 2 * int(ns.max())), rng)
    sid, soff, _ = gather(pilot)
    prec = ex["precision"][pilot].astype(np.uint8)
    te, td, _ = cpu_roundtrip(codec, sid, soff, prec, threads)
    rate = sid.size / (te + td)
    budget = int(max(0.01 * ids.numel(), rate * args.cpu_seconds))
    units = pick_sample_units(ns, budget, rng, always=[int(u) for u in longest])
    sid, soff, gdec = gather(units)
    prec = ex["precision"][units].astype(np.uint8)
    te, td, (heads, nwords, woff, words, cdec) = cpu_roundtrip(codec, sid, soff, prec, threads)
    bad = 0
    for i, u in enumerate(units):
        w0, w1 = int(ex["word_offsets"][u]), int(ex["word_offsets"][u + 1])
        cw = words[int(woff[i]): int(woff[i]) + int(nwords[i])]
        a, b = int(soff[i]), int(soff[i + 1])
        ok = (int(ex["heads"][u]) == int(heads[i]) and w1 - w0 == cw.size and np.array_equal(ex["words"][w0:w1], cw)
              and np.array_equal(gdec[a:b], cdec[a:b]))
        mx = int(sid[b - 1])
        ok = ok and int(prec[i]) == (mx - 1).bit_length()
        bad += 0 if ok else 1
    parity = {"checked": True, "units_checked": int(units.size), "ids_checked": int(sid.size),
              "fraction_of_ids": sid.size / ids.numel(), "includes_32_longest_units": True, "mismatching_units": bad,
              "bit_exact": bad == 0, "against": kind}
    cpu_base = {"value": sid.size / (te + td), "unit": "ids/s", "cores": threads, "kind": kind,
                "sample": f"{units.size} of {ns.size} ROC units ({sid.size} ids, uniform random units + the 32 longest), "
                          f"encode+decode once with {threads} OpenMP threads, schedule(dynamic)",
                "encode_ids_per_s": sid.size / te, "decode_ids_per_s": sid.size / td, "seconds": te + td}
    if bad:
        print(f"bench.py: PARITY FAILURE on {bad} units", file=sys.stderr)
    return parity, cpu_base


def ef_section(args, ctx, offsets, ids, dev, peak):
    import torch

    n_ids = int(ids.numel())
    enc_ms, dec_ms, meta_ms = [], [], []
    eb = None
    for it in range(args.warmup + args.steps):
        if eb is not None:
            eb.free()
        eb = ctx.ef_encode(offsets, ids, sorted_ids=True)
        be = dict(ctx.last_kernel_breakdown())
        out, _ = eb.decode(device=dev)
        bd = dict(ctx.last_kernel_breakdown())
        if it >= args.warmup:
            enc_ms.append(be["k_ef_encode"])
            meta_ms.append(be.get("k_unit_meta", 0.0))
            dec_ms.append(bd["k_ef_decode"])
    exact = bool(torch.equal(out, ids))
    comp = eb.bits_total / 8.0
    eb.free()
    del out
    e, d = float(np.mean(enc_ms)), float(np.mean(dec_ms))
    pm = float(np.mean(meta_ms))
    return {
        "bit_exact_roundtrip": exact, "bits_per_id": 8.0 * comp / n_ids,
        # prep = the metadata kernel in front of k_ef_encode (list ends for ascending input: the order / width check
        # of the ids happens inside k_ef_encode); frac_with_prep charges it to the encode
        "encode": {"kernel_ms": e, "prep_ms": pm, "ids_per_s": n_ids / (e * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + comp) / (e * 1e-3) / 1e9, "frac": (8.0 * n_ids + comp) / (e * 1e-3) / 1e9 / peak,
                   "frac_with_prep": (8.0 * n_ids + comp) / ((e + pm) * 1e-3) / 1e9 / peak},
        "decode": {"kernel_ms": d, "ids_per_s": n_ids / (d * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + comp) / (d * 1e-3) / 1e9, "frac": (8.0 * n_ids + comp) / (d * 1e-3) / 1e9 / peak},
    }


def wt_section(args, ctx, offsets, ids, sizes, dev, peak):
    """CompressedIDInvertedListsWaveletTree on the same lists (the workload's ids partition [0, N)): build, get_ids
    of every list, 10 M random get_single_id calls."""
    import torch

    n_ids = int(ids.numel())
    nlist = int(offsets.size - 1)
    enc_ms, dec_ms, sel_ms = [], [], []
    wb = None
    for it in range(2 + min(args.steps, 3)):
        if wb is not None:
            wb.free()
        wb = ctx.wt_encode(offsets, ids)
        if it >= 2:
            enc_ms.append(ctx.last_kernel_ms())
    struct_bytes = float(wb.bits_bytes + wb.aux_bytes)
    for it in range(2):
        out, _ = wb.decode(device=dev)
        dec_ms.append(ctx.last_kernel_ms())
    exact = bool(torch.equal(out, ids))
    del out
    g = torch.Generator(device=dev).manual_seed(11)
    nq = 10_000_000
    sz = torch.as_tensor(np.asarray(sizes, dtype=np.int64), device=dev)
    off = torch.as_tensor(offsets[:-1].astype(np.int64), device=dev)
    ql = torch.randint(0, nlist, (nq,), device=dev, generator=g)
    ql = ql[sz[ql] > 0]
    qo = torch.minimum((torch.rand(ql.numel(), device=dev, generator=g) * sz[ql]).long(), sz[ql] - 1)
    for it in range(2):
        got = wb.select(ql, qo, device=dev)
        sel_ms.append(ctx.last_kernel_ms())
    exact = exact and bool(torch.equal(got, ids[off[ql] + qo]))
    levels = wb.levels
    wb.free()
    e, d, q = float(np.mean(enc_ms)), dec_ms[-1], sel_ms[-1]
    return {
        "bit_exact_roundtrip": exact, "levels": levels, "bits_per_id": 8.0 * struct_bytes / n_ids,
        # algorithmic bytes: the ids in, the structure out (build) / the structure in, the ids out (decode)
        "encode": {"kernel_ms": e, "ids_per_s": n_ids / (e * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + struct_bytes) / (e * 1e-3) / 1e9,
                   "frac": (8.0 * n_ids + struct_bytes) / (e * 1e-3) / 1e9 / peak,
                   # what the level passes really stream: 4 B read for the bits + 4 B read and 4 B written by the partition
                   "pass_GBs": (8.0 * n_ids + 4.0 * n_ids * (3 * levels - 2) + struct_bytes) / (e * 1e-3) / 1e9},
        "decode": {"kernel_ms": d, "ids_per_s": n_ids / (d * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + struct_bytes) / (d * 1e-3) / 1e9,
                   "frac": (8.0 * n_ids + struct_bytes) / (d * 1e-3) / 1e9 / peak},
        "select": {"queries": int(ql.numel()), "kernel_ms": q, "queries_per_s": ql.numel() / (q * 1e-3)},
    }


def e2e_buffers(ids):
    import torch

    n_ids = int(ids.numel())
    host_in = torch.empty(n_ids, dtype=torch.int64, pin_memory=True)
    host_in.copy_(ids)
    host_out = torch.empty(n_ids, dtype=torch.int64, pin_memory=True)
    torch.cuda.synchronize()
    return host_in, host_out


def e2e_section(args, ctx, offsets, ids, world, dev, barrier, host_bufs):
    import torch
    import torch.distributed as dist

    n_ids = int(ids.numel())
    host_in, host_out = host_bufs
    hin, hout = host_in.numpy(), host_out.numpy()
    from vector_db_id_compression_b200 import capi

    def step():
        blob = ctx.roc_encode(offsets, hin, sorted_ids=True, max_unit=args.max_unit)  # H2D inside
        p, mem = capi._ptr(hout)
        off = np.zeros(blob.nlist + 1, np.uint64)
        capi._check(ctx._l.idc_roc_decode(ctx._h, blob._h, None, blob.nlist, p, 8, mem, off.ctypes.data))  # D2H inside
        nb = blob.ans_bytes
        blob.free()
        return nb

    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    barrier()
    t = time.perf_counter() - t0
    tt = torch.tensor([t], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    ok = bool(torch.equal(torch.sort(host_out[: int(offsets[1])])[0], host_in[: int(offsets[1])]))
    return {"value": n_ids * world * args.e2e_steps / t, "unit": "ids/s", "h2d_bytes_per_step": 8 * n_ids,
            "d2h_bytes_per_step": 8 * n_ids, "steps": args.e2e_steps, "ms_per_step": 1e3 * t / args.e2e_steps,
            "first_list_roundtrip_ok": ok,
            "path": "idc_roc_encode(IDC_MEM_HOST, pinned) -> idc_roc_decode(IDC_MEM_HOST, pinned)"}


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout: keep the real stdout for it and send everything else that a
    # library may print there (e.g. "NCCL version ...") to stderr.
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
