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
    ap.add_argument("--parity-frac", type=float, default=0.25, help="share of the ids checked bit-exact against the CPU reference")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = --steps")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e: one step at a time only")
    ap.add_argument("--no-ef", action="store_true")
    ap.add_argument("--no-wt", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1-C4 / control / s=0.5 sub-objects (N=1 only)")
    ap.add_argument("--no-accessors", action="store_true", help="skip the per-call accessor leg (N=1 only)")
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
    from vector_db_id_compression_b200.sharding import unit_table as ut

    _, starts, ns = ut(np.asarray(offsets).astype(np.int64), max_unit)
    return starts, ns


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
    m = np.maximum(max_ids.astype(np.int64) - 1, 0).astype(np.uint64)
    out = np.zeros(m.size, dtype=np.uint8)
    for _ in range(64):
        nz = m > 0
        if not nz.any():
            break
        out[nz] += 1
        m[nz] >>= np.uint64(1)
    return out


def gather_units(ids, out, starts, ns, units, dev):
    """ids (and decoded ids) of the given units, concatenated; vectorised index arithmetic on the device."""
    import torch

    n_u = torch.as_tensor(ns[units], device=dev)
    s_u = torch.as_tensor(starts[units], device=dev)
    soff = np.zeros(units.size + 1, dtype=np.uint64)
    soff[1:] = np.cumsum(ns[units])
    total = int(soff[-1])
    base = torch.repeat_interleave(s_u - torch.as_tensor(soff[:-1].astype(np.int64), device=dev), n_u)
    idx = base + torch.arange(total, device=dev, dtype=torch.int64)
    sid = ids[idx].cpu().numpy().astype(np.uint64)
    dec = out[idx].cpu().numpy().astype(np.uint64) if out is not None else None
    return sid, soff, dec


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
        sid, soff, _ = gather_units(ids, None, starts, ns, units, dev)
        prec = precision_rule_np(sid[soff[1:].astype(np.int64) - 1])
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
        "l2_fetch_granularity": os.environ.get("IDC_L2_FETCH", "driver default") + " B (IDC_L2_FETCH: opt-in, device-global, restored on ctx destroy)",
        "parallelism": "value / e2e: one workload per GPU (weak scaling, no data-path collective); "
                       "sharded: ONE index owned by rank 0, NCCL scatter of id blocks + gather of blobs (strong scaling)",
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

    # The ROC kernels read isolated 32-byte sectors: the bench opts into the 32-byte L2 fetch granularity (a
    # device-global limit, so the library leaves it alone unless asked; restored when the context is destroyed).
    os.environ.setdefault("IDC_L2_FETCH", "32")
    from vector_db_id_compression_b200.capi import Context

    sizes, offsets, ids = make_workload(args, dev, args.seed + 7919 * rank)
    n_ids = int(ids.numel())
    ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
    ctx.set_timing(True)

    # ---------------- device-resident timed region (value)
    t_leg = time.perf_counter()
    timed = time_roc(args, ctx, offsets, ids, dev, args.steps, args.warmup, world, barrier, sample_clocks=local)
    blob, out = timed["blob"], timed["out"]
    ms_max = timed["ms_total"]
    value = n_ids * world * args.steps / (ms_max * 1e-3)
    avg = timed["kernel_ms"]
    legs = {"value": time.perf_counter() - t_leg}

    # ---------------- parity: sampled units (+ the longest) against the CPU reference; also the cpu_baseline
    info = dict(nunits=blob.nunits, ans_bytes=blob.ans_bytes, total_words=blob.total_words)
    parity = cpu_base = None
    t_leg = time.perf_counter()
    if rank == 0 and not args.no_cpu_baseline:
        parity, cpu_base = roc_parity(args, blob, out, offsets, ids, dev, frac=args.parity_frac, always_longest=32,
                                      cpu_seconds=max(args.cpu_seconds, 60.0), one_thread=True)  # (the pilot underestimates the rate: the cap that binds is parity_frac)
    elif rank == 0:
        parity = {"checked": False}
    legs["parity"] = time.perf_counter() - t_leg

    # ---------------- roofline of the dominant kernel (algorithmic bytes / event time)
    peak, peak_src = measured_peak_gbs()
    roofline = roc_roofline(avg, n_ids, blob.ans_bytes, peak, peak_src, args)
    blob.free()
    del out

    # ---------------- Elias-Fano (the HBM-bound codec) on the same lists
    ef = None
    if not args.no_ef:
        t_leg = time.perf_counter()
        ef = ef_section(args, ctx, offsets, ids, dev, peak, cpu=(rank == 0 and not args.no_cpu_baseline))
        legs["ef"] = time.perf_counter() - t_leg

    # ---------------- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # every rank pins 16 GB of host memory for this leg; if the box cannot give that to all ranks the leg is
        # reported as failed (on every rank alike, so that the collectives inside stay matched) instead of taking
        # the device-resident numbers down with it
        t_leg = time.perf_counter()
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
        del host_bufs
        legs["e2e"] = time.perf_counter() - t_leg

    # ---------------- the north-star multi-GPU path: ONE index owned by rank 0, NCCL scatter / encode / gather
    sharded = None
    if not args.no_sharded:
        t_leg = time.perf_counter()
        try:
            sharded = sharded_section(args, ctx, offsets, ids, world, rank, dev, barrier)
        except Exception as ex:  # noqa: BLE001
            if world > 1:
                raise  # a rank that left the collectives would hang the others: fail loudly instead
            sharded = {"error": str(ex)}
        legs["sharded"] = time.perf_counter() - t_leg

    # ---------------- wavelet tree (the third id index of the plugin surface) on the same lists
    wt = None
    if not args.no_wt:
        t_leg = time.perf_counter()
        try:
            wt = wt_section(args, ctx, offsets, ids, sizes, dev, peak)
        except Exception as ex:  # noqa: BLE001 -- never takes the headline numbers down with it
            wt = {"error": str(ex)}
        legs["wt"] = time.perf_counter() - t_leg
    del ids
    torch.cuda.empty_cache()

    # ---------------- the other configurations of BASELINE.json, the control and s = 0.5 (one GPU, rank 0)
    configs = accessors = None
    if world == 1 and not args.no_configs:
        t_leg = time.perf_counter()
        configs = configs_section(args, ctx, dev, peak)
        legs["configs"] = time.perf_counter() - t_leg
    if world == 1 and not args.no_accessors:
        t_leg = time.perf_counter()
        try:
            accessors = accessors_section(args)
        except Exception as ex:  # noqa: BLE001
            accessors = {"error": str(ex)}
        legs["accessors"] = time.perf_counter() - t_leg

    if rank == 0:
        line = {
            "metric": "ROC encode+decode ids/s (bit-exact round trip)", "value": value, "unit": "ids/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, sizes),
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(timed["launches"]),
            "clocks": timed["clocks"], "parity": parity,
            "roc": {"encode_ids_per_s": n_ids / (sum(avg.get(k, 0) for k in ("k_unit_meta", "k_enc_records", "k_roc_encode", "k_roc_compact")) * 1e-3),
                    "decode_ids_per_s": n_ids / (sum(avg.get(k, 0) for k in ("memset_ws", "k_roc_decode")) * 1e-3),
                    "bits_per_id": 8.0 * info["ans_bytes"] / n_ids, "units": info["nunits"],
                    "wall_ms_per_step": 1e3 * timed["wall_s"] / args.steps},
            "ef": ef, "wt": wt, "sharded": sharded, "configs": configs, "accessors": accessors,
            "leg_seconds": {k: round(v, 2) for k, v in legs.items()},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_roc(args, ctx, offsets, ids, dev, steps, warmup, world, barrier, sample_clocks=None, max_unit=None):
    """W warm-up + K timed steps of ROC encode-everything + decode-everything with the ids resident in HBM; CUDA
    events on the codec's stream, max over ranks. Returns the last blob / output for the parity check."""
    import torch
    import torch.distributed as dist

    mu = max_unit or args.max_unit
    kern_ms = {}

    def one_step(collect: bool):
        blob = ctx.roc_encode(offsets, ids, sorted_ids=True, max_unit=mu)
        if collect:
            for k, v in ctx.last_kernel_breakdown():
                kern_ms.setdefault(k, []).append(v)
        out, _ = blob.decode(device=dev)
        if collect:
            for k, v in ctx.last_kernel_breakdown():
                kern_ms.setdefault(k, []).append(v)
        return blob, out

    blob = out = None
    for _ in range(warmup):
        if blob is not None:
            blob.free()
        blob, out = one_step(False)
    sampler = ClockSampler(sample_clocks) if sample_clocks is not None else None
    barrier()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        if blob is not None:
            blob.free()
        del out
        blob, out = one_step(True)
    ev1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    t_dev = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    return {"blob": blob, "out": out, "ms_total": float(t_dev.item()), "wall_s": t_wall, "clocks": clocks,
            "launches": ctx.launch_count - launches0, "kernel_ms": {k: float(np.mean(v)) for k, v in kern_ms.items()}}


def roc_roofline(avg, n_ids, ans_bytes, peak, peak_src, args, traffic_ok=True):
    enc_bytes = 8.0 * n_ids + ans_bytes          # read int64 ids, write streams
    dec_bytes = ans_bytes + 8.0 * n_ids          # read streams, write int64 ids
    dom = max(("k_roc_encode", "k_roc_decode"), key=lambda k: avg.get(k, 0.0))
    dom_bytes = enc_bytes if dom == "k_roc_encode" else dec_bytes
    achieved = dom_bytes / (avg[dom] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(dom, n_ids, args.zipf_s) if traffic_ok else (None, None)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": avg[dom],
                "other": {k: {"ms": avg[k]} for k in avg if k != dom}}
    for k, b in (("k_roc_encode", enc_bytes), ("k_roc_decode", dec_bytes)):
        if k in avg and k != dom:
            roofline["other"][k].update(achieved=b / (avg[k] * 1e-3) / 1e9, frac=b / (avg[k] * 1e-3) / 1e9 / peak)
    return roofline


def roc_parity(args, blob, out, offsets, ids, dev, frac, always_longest=0, cpu_seconds=40.0, one_thread=False, max_unit=None):
    """Bit-exact check of `frac` of the ids (whole units, uniform random + the longest) against the CPU codec:
    (head, words) of every checked unit byte-equal, precision per the reference rule, decoded array equal INCLUDING
    order. The CPU codec is timed on the way (cpu_baseline)."""
    codec, kind = cpu_codec()
    threads = os.cpu_count() or 1
    ex = blob.export()
    starts, ns = unit_table(offsets, max_unit or args.max_unit)
    assert ns.size == blob.nunits
    rng = np.random.default_rng(99)
    n_total = int(ns.sum())
    longest = np.argsort(-ns, kind="stable")[:always_longest]
    if frac >= 1.0:
        units = np.nonzero(ns > 0)[0]
    else:
        # pilot -> what the CPU can do in cpu_seconds caps the sample
        pilot = pick_sample_units(ns, max(200_000, 2 * int(ns.max())), rng)
        sid, soff, _ = gather_units(ids, None, starts, ns, pilot, dev)
        te, td, _ = cpu_roundtrip(codec, sid, soff, ex["precision"][pilot].astype(np.uint8), threads)
        rate = sid.size / (te + td)
        budget = int(min(frac * n_total, max(0.01 * n_total, rate * cpu_seconds)))
        units = pick_sample_units(ns, budget, rng, always=[int(u) for u in longest])
    sid, soff, gdec = gather_units(ids, out, starts, ns, units, dev)
    prec = ex["precision"][units].astype(np.uint8)
    te, td, (heads, nwords, woff, words, cdec) = cpu_roundtrip(codec, sid, soff, prec, threads)
    # vectorised comparison: heads, word counts, precisions per unit; words and decoded ids as flat arrays
    wo = ex["word_offsets"].astype(np.int64)
    g_nw = (wo[units + 1] - wo[units])
    c_nw = np.asarray(nwords[: units.size]).astype(np.int64)
    unit_ok = (ex["heads"][units] == np.asarray(heads[: units.size])) & (g_nw == c_nw)
    last = sid[soff[1:].astype(np.int64) - 1]
    unit_ok &= prec == precision_rule_np(last)
    both = np.nonzero(g_nw == c_nw)[0]
    if both.size:
        n_w = g_nw[both]
        tot = int(n_w.sum())
        rel = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(n_w) - n_w, n_w)
        gw = ex["words"][np.repeat(wo[units[both]], n_w) + rel]
        cw = words[np.repeat(np.asarray(woff[:-1]).astype(np.int64)[both], n_w) + rel]
        neq = gw != cw
        if neq.any():
            unit_ok[both[np.unique(np.repeat(np.arange(both.size), n_w)[neq])]] = False
    dneq = gdec != cdec[: gdec.size]
    if dneq.any():
        unit_ok[np.unique(np.repeat(np.arange(units.size), ns[units])[dneq])] = False
    bad = int((~unit_ok).sum())
    parity = {"checked": True, "units_checked": int(units.size), "ids_checked": int(sid.size),
              "fraction_of_ids": sid.size / max(n_total, 1), "includes_longest_units": int(always_longest),
              "mismatching_units": bad, "bit_exact": bad == 0, "against": kind,
              "what": "(head, words) byte-equal, precision rule, decoded ids equal including order"}
    cpu_base = {"value": sid.size / (te + td), "unit": "ids/s", "cores": threads, "kind": kind,
                "sample": f"{units.size} of {ns.size} ROC units ({sid.size} ids, uniform random units"
                          f"{' + the %d longest' % always_longest if always_longest else ''}), "
                          f"encode+decode once with {threads} OpenMP threads, schedule(dynamic)",
                "encode_ids_per_s": sid.size / te, "decode_ids_per_s": sid.size / td, "seconds": te + td}
    if one_thread:
        small = pick_sample_units(ns, 1_500_000, np.random.default_rng(7))
        s1, o1, _ = gather_units(ids, None, starts, ns, small, dev)
        t1e, t1d, _ = cpu_roundtrip(codec, s1, o1, ex["precision"][small].astype(np.uint8), 1)
        cpu_base["one_thread"] = {"value": s1.size / (t1e + t1d), "unit": "ids/s", "cores": 1, "ids": int(s1.size),
                                  "encode_ids_per_s": s1.size / t1e, "decode_ids_per_s": s1.size / t1d}
    if bad:
        print(f"bench.py: PARITY FAILURE on {bad} units", file=sys.stderr)
    return parity, cpu_base


def ef_section(args, ctx, offsets, ids, dev, peak, cpu=False, steps=None, warmup=None):
    import torch

    n_ids = int(ids.numel())
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    enc_ms, dec_ms, meta_ms = [], [], []
    eb = None
    for it in range(warmup + steps):
        if eb is not None:
            eb.free()
        eb = ctx.ef_encode(offsets, ids, sorted_ids=True)
        be = dict(ctx.last_kernel_breakdown())
        out, _ = eb.decode(device=dev)
        bd = dict(ctx.last_kernel_breakdown())
        if it >= warmup:
            enc_ms.append(be["k_ef_encode"])
            meta_ms.append(be.get("k_unit_meta", 0.0) + be.get("k_ef_tile_desc", 0.0) + be.get("k_ef_finish_chunks", 0.0))
            dec_ms.append(bd["k_ef_decode"])
    exact = bool(torch.equal(out, ids))
    comp = eb.bits_total / 8.0
    del out
    e, d = float(np.mean(enc_ms)), float(np.mean(dec_ms))
    pm = float(np.mean(meta_ms))
    res = {
        "bit_exact_roundtrip": exact, "bits_per_id": 8.0 * comp / max(n_ids, 1),
        # prep = the small kernels around k_ef_encode (list ends for ascending input -- the order / width check of the
        # ids happens inside k_ef_encode --, the tile descriptors, the chunk-count fix-up); frac_with_prep charges them
        # to the encode
        "encode": {"kernel_ms": e, "prep_ms": pm, "ids_per_s": n_ids / (e * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + comp) / (e * 1e-3) / 1e9, "frac": (8.0 * n_ids + comp) / (e * 1e-3) / 1e9 / peak,
                   "frac_with_prep": (8.0 * n_ids + comp) / ((e + pm) * 1e-3) / 1e9 / peak},
        "decode": {"kernel_ms": d, "ids_per_s": n_ids / (d * 1e-3),
                   "achieved_GBs": (8.0 * n_ids + comp) / (d * 1e-3) / 1e9, "frac": (8.0 * n_ids + comp) / (d * 1e-3) / 1e9 / peak},
    }
    if cpu:
        res["cpu_baseline"], res["words_vs_restatement"] = ef_cpu_baseline(eb, offsets, ids, dev)
    eb.free()
    return res


def ef_cpu_baseline(eb, offsets, ids, dev, budget_ids=60_000_000):
    """The Elias-Fano restatement (oracle/ef_oracle.c: the reference class needs ot/succinct, absent) on a sample of
    whole lists: OpenMP over lists like custom_invlists_impl.cpp:234 and one thread; its words are compared with the
    GPU blob's on the way."""
    import oracle

    threads = os.cpu_count() or 1
    off = np.asarray(offsets).astype(np.int64)
    sizes = np.diff(off)
    rng = np.random.default_rng(5)
    # whole lists, none longer than a quarter of the budget (one 8.6e7-id list would be the whole sample)
    cand = np.where(sizes <= budget_ids // 4, sizes, 0)
    lists = pick_sample_units(cand, min(budget_ids, int(cand.sum())), rng)
    sid, soff, _ = gather_units(ids, None, off[:-1], sizes, lists, dev)
    t0 = time.perf_counter()
    enc = oracle.ef.encode_lists(soff, sid, nthreads=threads)
    t1 = time.perf_counter()
    dec = oracle.ef.decode_lists(soff, enc, nthreads=threads)
    t2 = time.perf_counter()
    small = lists[: max(1, lists.size // 16)]
    s1, o1, _ = gather_units(ids, None, off[:-1], sizes, small, dev)
    t3 = time.perf_counter()
    e1 = oracle.ef.encode_lists(o1, s1, nthreads=1)
    t4 = time.perf_counter()
    oracle.ef.decode_lists(o1, e1, nthreads=1)
    t5 = time.perf_counter()
    # bit level: the blob's words of the sampled lists against the restatement's
    ex = eb.export()
    ok = bool(np.array_equal(dec, sid)) and bool(np.array_equal(ex["l"][lists], enc["l"][: lists.size]))
    for name in ("low", "high"):
        go = ex[name + "_offsets"].astype(np.int64)
        n_w = np.diff(enc[name + "_off"].astype(np.int64))
        ok = ok and bool(np.array_equal(go[lists + 1] - go[lists], n_w))
        if not ok:
            break
        tot = int(n_w.sum())
        rel = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(n_w) - n_w, n_w)
        ok = ok and bool(np.array_equal(ex[name][np.repeat(go[lists], n_w) + rel], enc[name][:tot]))
    base = {"value": sid.size / (t2 - t0), "unit": "ids/s", "cores": threads, "kind": "port",
            "sample": f"{lists.size} whole lists ({sid.size} ids), OpenMP over lists, {threads} threads",
            "encode_ids_per_s": sid.size / (t1 - t0), "decode_ids_per_s": sid.size / (t2 - t1),
            "one_thread": {"encode_ids_per_s": s1.size / (t4 - t3), "decode_ids_per_s": s1.size / (t5 - t4), "ids": int(s1.size)}}
    return base, {"lists": int(lists.size), "ids": int(sid.size), "bit_equal": ok}


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
    # wt_type = 1 (rrr_vector<63> flavour): the same index with block-compressed levels
    rrr = None
    try:
        rb = None
        r_enc = []
        for it in range(2):
            if rb is not None:
                rb.free()
            rb = ctx.wt_encode(offsets, ids, wt_type=1)
            r_enc.append(ctx.last_kernel_ms())
        comp_ms = max(0.0, r_enc[-1] - e)  # what the block compression adds to the plain build (kernel milliseconds of the call)
        r_bytes = float(rb.bits_bytes + rb.aux_bytes)
        r_sel = []
        for it in range(2):
            got1 = rb.select(ql, qo, device=dev)
            r_sel.append(ctx.last_kernel_ms())
        r_dec = []
        for it in range(2):
            out1, _ = rb.decode(device=dev)
            r_dec.append(ctx.last_kernel_ms())
        rrr = {"bits_per_id": 8.0 * r_bytes / n_ids, "build_ms": r_enc[-1], "compress_ms": comp_ms, "select_ms": r_sel[-1],
               "queries_per_s": ql.numel() / (r_sel[-1] * 1e-3), "decode_ms": r_dec[-1],
               "exact": bool(torch.equal(got1, got)) and bool(torch.equal(out1, ids))}
        del out1, got1
        rb.free()
    except Exception as ex:  # a failure here must not take the plain numbers down
        rrr = {"error": str(ex)[:200]}
    return {
        "wt_type_1": rrr,
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


def pcie_ceiling(host_in, host_out, world, dev, barrier, chunk_bytes=1 << 30, reps=4):
    """What the host link gives this rank while ALL ranks copy at the same time: pinned-host -> device alone, device ->
    pinned-host alone, and both directions at once (two streams). The e2e step moves 8 B/id up and 8 B/id down, the
    download after the upload, so its copy floor is bytes / h2d + bytes / d2h."""
    import torch
    import torch.distributed as dist

    n = min(chunk_bytes // 8, int(host_in.numel()))
    dbuf_a = torch.empty(n, dtype=torch.int64, device=dev)
    dbuf_b = torch.empty(n, dtype=torch.int64, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    nchunk = max(1, min(reps, int(host_in.numel()) // n))

    def run(up, down):
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for k in range(nchunk):
            if up:
                with torch.cuda.stream(s1):
                    dbuf_a.copy_(host_in[k * n:(k + 1) * n], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    host_out[k * n:(k + 1) * n].copy_(dbuf_b, non_blocking=True)
        s1.synchronize()
        s2.synchronize()
        return 8.0 * n * nchunk / (time.perf_counter() - t0) / 1e9  # GB/s per direction

    run(True, True)
    vals = [run(True, False), run(False, True), run(True, True)]
    out = {}
    for name, v in zip(("h2d_GBs", "d2h_GBs", "duplex_GBs_per_direction"), vals):
        allr = [v]
        if world > 1:
            g = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(g, torch.tensor([v], device=dev, dtype=torch.float64))
            allr = [float(x.item()) for x in g]
        out[name] = [round(x, 2) for x in allr]
    del dbuf_a, dbuf_b
    return out


def e2e_pipelined(args, offsets, hin, hout, steps, dev, barrier, world):
    """Two steps in flight through the same two C-ABI calls (see e2e_section). Returns ms_per_step (max over ranks)."""
    import queue
    import threading

    import torch
    import torch.distributed as dist

    from vector_db_id_compression_b200 import capi

    ctx_e, ctx_d = capi.Context(dev.index or 0), capi.Context(dev.index or 0)
    q = queue.Queue(maxsize=1)
    err = []

    def encoder(k):
        try:
            for _ in range(k):
                blob = ctx_e.roc_encode(offsets, hin, sorted_ids=True, max_unit=args.max_unit)  # H2D inside
                # a blob belongs to its context (pool, lock): it crosses to the decoder's context in its wire form,
                # device to device (idc_roc_blob_export_payload -> idc_roc_blob_assemble, what the NCCL gather uses too)
                payload = blob.export_payload(device=dev)
                blob.free()
                q.put(payload)
        except Exception as ex:  # noqa: BLE001
            err.append(ex)
        finally:
            q.put(None)

    def decoder():
        p, mem = capi._ptr(hout)
        while True:
            payload = q.get()
            if payload is None:
                return
            blob = None
            try:
                blob = ctx_d.roc_assemble(offsets, payload, max_unit=args.max_unit)
                off = np.zeros(blob.nlist + 1, np.uint64)
                capi._check(ctx_d._l.idc_roc_decode(ctx_d._h, blob._h, None, blob.nlist, p, 8, mem, off.ctypes.data))  # D2H inside
            except Exception as ex:  # noqa: BLE001
                err.append(ex)
            finally:
                if blob is not None:
                    blob.free()

    def run(k):
        te, td = threading.Thread(target=encoder, args=(k,)), threading.Thread(target=decoder)
        te.start()
        td.start()
        te.join()
        td.join()

    run(2)  # warm both contexts' workspaces
    barrier()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    barrier()
    t_local = time.perf_counter() - t0
    ctx_e.close()
    ctx_d.close()
    if err:
        raise err[0]
    tt = torch.tensor([t_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return {"ms_per_step": 1e3 * float(tt.item()) / steps, "t_local": t_local}


def e2e_section(args, ctx, offsets, ids, world, dev, barrier, host_bufs):
    import torch
    import torch.distributed as dist

    n_ids = int(ids.numel())
    steps = args.e2e_steps or args.steps
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
    for _ in range(steps):
        step()
    barrier()
    t_local = time.perf_counter() - t0
    # the same steps, two in flight: step k's decode + download run while step k + 1 uploads and encodes (PCIe is full
    # duplex and one step's kernels leave most issue slots idle). Two host threads, each with its own context (stream,
    # workspaces); the blob crosses from one to the other in its wire form, device to device -- what a server that
    # compresses and serves batches back to back does with the same C-ABI calls. Same timed region: every step's 8 GB
    # up and 8 GB down are inside.
    piped = None
    if not getattr(args, "no_e2e_pipeline", False):
        try:
            piped = e2e_pipelined(args, offsets, hin, hout, steps, dev, barrier, world)
        except Exception as ex:
            piped = {"error": str(ex)[:200]}
    tt = torch.tensor([t_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t = float(tt.item())
    # the whole round trip: every decoded list, sorted, equals its input list (a checksum of the sorted output per
    # list would hide nothing more: the comparison runs on the GPU over all ids of this rank)
    back = host_out.to(dev, non_blocking=False)
    lab = torch.repeat_interleave(torch.arange(offsets.size - 1, device=dev, dtype=torch.int64),
                                  torch.as_tensor(np.diff(offsets.astype(np.int64)), device=dev))
    key, _ = torch.sort(back + (lab << 32))
    ok = bool(torch.equal(key - (lab << 32), ids))
    del back, lab, key
    # per-rank PCIe rates: 8 B/id up during the encode, 8 B/id down during the decode, overlapped with the kernels
    gbs = 16.0 * n_ids / t_local / 1e9 * steps
    rates = [gbs]
    if world > 1:
        allr = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allr, torch.tensor([gbs], device=dev, dtype=torch.float64))
        rates = [float(x.item()) for x in allr]
    res = {"value": n_ids * world * steps / t, "unit": "ids/s", "h2d_bytes_per_step": 8 * n_ids,
           "d2h_bytes_per_step": 8 * n_ids, "steps": steps, "ms_per_step": 1e3 * t / steps,
           "roundtrip_all_lists_ok": ok, "pcie_GBs_per_rank": [round(r, 2) for r in rates],
           "path": "idc_roc_encode(IDC_MEM_HOST, pinned) -> idc_roc_decode(IDC_MEM_HOST, pinned)"}
    if piped and "ms_per_step" in piped:
        # reported beside the headline, not as it: `value` stays the plain call sequence a user makes, one step at a time
        res["pipelined"] = {"value": n_ids * world / (piped["ms_per_step"] * 1e-3), "unit": "ids/s", "ms_per_step": piped["ms_per_step"],
                            "what": "two steps in flight: step k + 1 uploads and encodes on one host thread / context while step k "
                                    "is decoded and downloaded on another; the blob crosses in its wire form, device to device "
                                    "(idc_roc_blob_export_payload -> idc_roc_blob_assemble). The two logical ROC kernels do not fit "
                                    "the SMs' shared memory side by side, the copies do overlap"}
    elif piped:
        res["pipelined"] = piped
    # stated variant: the same round trip with 4-byte ids on the wire (ids < 2^31 here; faiss::idx_t is 8 bytes, so this is
    # NOT the headline): half the PCIe bytes
    try:
        if world == 1 and int(ids.max().item()) < (1 << 31):  # (N = 1 only: another 8 GB of pinned host memory per rank)
            h32 = torch.empty(n_ids, dtype=torch.int32, pin_memory=True)
            h32.copy_(ids.to(torch.int32))
            o32 = torch.empty(n_ids, dtype=torch.int32, pin_memory=True)
            hin32, hout32 = h32.numpy(), o32.numpy()

            def step32():
                blob = ctx.roc_encode(offsets, hin32, sorted_ids=True, max_unit=args.max_unit)
                p, mem = capi._ptr(hout32)
                off = np.zeros(blob.nlist + 1, np.uint64)
                capi._check(ctx._l.idc_roc_decode(ctx._h, blob._h, None, blob.nlist, p, 4, mem, off.ctypes.data))
                blob.free()

            step32()
            barrier()
            t0 = time.perf_counter()
            k32 = min(steps, 3)
            for _ in range(k32):
                step32()
            barrier()
            t32 = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t32, op=dist.ReduceOp.MAX)
            back32 = o32.to(dev).to(torch.int64)
            lab32 = torch.repeat_interleave(torch.arange(offsets.size - 1, device=dev, dtype=torch.int64),
                                            torch.as_tensor(np.diff(offsets.astype(np.int64)), device=dev))
            key32, _ = torch.sort(back32 + (lab32 << 32))
            res["id_bytes_4"] = {"value": n_ids * world * k32 / float(t32.item()), "unit": "ids/s", "ms_per_step": 1e3 * float(t32.item()) / k32,
                                 "h2d_bytes_per_step": 4 * n_ids, "d2h_bytes_per_step": 4 * n_ids, "steps": k32,
                                 "roundtrip_all_lists_ok": bool(torch.equal(key32 - (lab32 << 32), ids)),
                                 "what": "stated variant, not the headline: int32 ids in and out (id_bytes = 4 of the C ABI)"}
            del back32, lab32, key32, h32, o32
    except Exception as ex:
        res["id_bytes_4"] = {"error": str(ex)[:200]}
    try:
        ceil = pcie_ceiling(host_in, host_out, world, dev, barrier)
        # copy floor of one step on the slowest rank: the download can only follow the upload
        floor_ms = max(1e3 * (8.0 * n_ids / 1e9 / h + 8.0 * n_ids / 1e9 / d) for h, d in zip(ceil["h2d_GBs"], ceil["d2h_GBs"]))
        res["pcie_ceiling"] = dict(ceil, what="pinned host <-> device copies of 1 GiB pieces, all ranks at the same time",
                                   copy_floor_ms_per_step=round(floor_ms, 1), step_over_copy_floor=round(res["ms_per_step"] / floor_ms, 3))
    except Exception as ex:  # the measurement must never take the e2e number down with it
        res["pcie_ceiling"] = {"error": str(ex)[:200]}
    return res


# ------------------------------------------------------------------ sharded (rank 0 owns the index)

def sharded_section(args, ctx, offsets, ids, world, rank, dev, barrier):
    """BASELINE.json C5 / C4 as the north star states them: ONE index owned by rank 0; NCCL over NVLink scatters the
    raw id blocks and gathers the compressed blobs, every rank encodes (and decodes) its share. Device buffers end to
    end. Reported next to the 1-GPU encode of the same index measured in the same run on rank 0."""
    import torch
    import torch.distributed as dist

    from vector_db_id_compression_b200 import sharding, workloads as W

    codec = sharding.RocCudaCodec(ctx, args.max_unit)
    res = {"n_ranks": world, "what": "rank 0 owns the index: tensor broadcast of the offsets, contiguous-unit-range "
           "plan, ncclSend/Recv of id slices, per-rank idc_roc_encode, gather-v of device payloads, idc_roc_blob_assemble"}

    def maxr(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run(name, off, idt, reps):
        n = int(off[-1] - off[0]) if rank == 0 else 0
        nt = torch.tensor([n], device=dev, dtype=torch.int64)
        if world > 1:
            dist.broadcast(nt, src=0)
        n = int(nt.item())
        # -- whole step, CUDA events on the current stream (NCCL waits are stream-ordered), max over ranks
        step_ms, whole, local_blob, plan = [], None, None, None
        for it in range(1 + reps):
            if whole is not None:
                whole.free()
            if local_blob is not None:
                local_blob.free()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            whole, local_blob, plan = sharding.encode_sharded(off, idt, codec, dev)
            e1.record()
            torch.cuda.synchronize()
            if it:
                step_ms.append(maxr(e0.elapsed_time(e1)))
        # -- phases (each closed by a device sync; max over ranks per phase)
        marks = []

        def tick(nm):
            # every phase ends at a rendezvous of all ranks: a rank that finished its encode early would otherwise book
            # the wait for the slowest rank under "gather"
            torch.cuda.synchronize()
            barrier()
            marks.append((nm, time.perf_counter()))

        if whole is not None:
            whole.free()
        local_blob.free()
        barrier()
        t0 = time.perf_counter()
        whole, local_blob, plan = sharding.encode_sharded(off, idt, codec, dev, timer=tick)
        phases, prev = {}, t0
        for nm, t in marks:
            phases[nm] = maxr(1e3 * (t - prev))
            prev = t
        # -- ids already resident per rank (no scatter)
        mine = sharding.scatter_id_blocks(plan, idt, dev) if world > 1 else idt[int(plan["ecut"][0]): int(plan["ecut"][1])]
        res_ms = []
        for it in range(1 + reps):
            if whole is not None:
                whole.free()
            local_blob.free()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            whole, local_blob, plan = sharding.encode_sharded(off, idt, codec, dev, resident_ids=mine)
            e1.record()
            torch.cuda.synchronize()
            if it:
                res_ms.append(maxr(e0.elapsed_time(e1)))
        # -- every rank decodes its own units; round trip against its id block
        dec_times = []
        for it in range(2):  # the first pass allocates the decoder's workspaces: the second one is the measurement
            dec = None
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec, _ = local_blob.decode(device=dev)
            e1.record()
            torch.cuda.synchronize()
            dec_times.append(maxr(e0.elapsed_time(e1)))
        dec_ms = dec_times[-1]
        lo = plan["local_offsets"][rank if world > 1 else 0].astype(np.int64)
        lab = torch.repeat_interleave(torch.arange(lo.size - 1, device=dev, dtype=torch.int64), torch.as_tensor(np.diff(lo), device=dev))
        srt, _ = torch.sort(dec[: mine.numel()] + (lab << 32))
        rt_ok = torch.tensor([1 if bool(torch.equal(srt - (lab << 32), mine)) else 0], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(rt_ok, op=dist.ReduceOp.MIN)
        del dec, lab, srt
        out = {"n_ids": n, "step_ms": float(np.mean(step_ms)), "phases_ms": {k: round(v, 3) for k, v in phases.items()},
               "resident_step_ms": float(np.mean(res_ms)), "decode_ms": dec_ms, "roundtrip_ok": bool(rt_ok.item()),
               "ids_per_s_encode": n / (np.mean(step_ms) * 1e-3), "ids_per_s_encode_decode": n / ((np.mean(step_ms) + dec_ms) * 1e-3)}
        # -- rank 0: the same index on ONE GPU, same run; assembled blob byte-identical?
        if rank == 0:
            one_ms = []
            single = None
            for it in range(1 + reps):
                if single is not None:
                    single.free()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                single = ctx.roc_encode(off, idt, sorted_ids=True, max_unit=args.max_unit)
                e1.record()
                torch.cuda.synchronize()
                if it:
                    one_ms.append(e0.elapsed_time(e1))
            a, b = whole.export_payload(device=dev), single.export_payload(device=dev)
            same = all(bool(torch.equal(a[k], b[k])) for k in a) and whole.ans_bytes == single.ans_bytes
            for it in range(2):  # the second pass is the measurement (the first allocates workspaces)
                d1 = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                d1, _ = single.decode(device=dev)
                e1.record()
                torch.cuda.synchronize()
                one_dec = e0.elapsed_time(e1)
            del d1, a, b
            t1 = float(np.mean(one_ms))
            out.update(byte_identical_to_one_gpu=bool(same), one_gpu_encode_ms=t1, one_gpu_decode_ms=one_dec,
                       blob_bytes=int(single.ans_bytes),
                       strong_efficiency=t1 / (world * out["step_ms"]),
                       strong_efficiency_resident=t1 / (world * out["resident_step_ms"]),
                       strong_efficiency_encode_decode=(t1 + one_dec) / (world * (out["step_ms"] + dec_ms)))
            if world > 1:
                sent = 8.0 * n * (world - 1) / world
                out["scatter_GBs"] = sent / (phases["scatter"] * 1e-3) / 1e9 if phases.get("scatter") else None
                out["gather_GBs"] = single.ans_bytes * (world - 1) / world / (phases["gather"] * 1e-3) / 1e9 if phases.get("gather") else None
                enc_share = phases.get("encode", 0.0) / max(sum(phases.values()), 1e-9)
                out["limiter"] = ("codec kernels: the serial chains of the longest units (up to max_unit steps of ~1.5 us) do not get "
                                  "shorter when the units are spread over more GPUs" if enc_share > 0.6 else
                                  "rank 0's NVLink egress (scatter) / ingress (gather) and host-side planning")
            single.free()
        if world > 1:
            barrier()
        whole and whole.free()
        local_blob.free()
        del mine
        return out

    res["c5"] = run("c5", offsets if rank == 0 else None, ids if rank == 0 else None, reps=2)
    c4_off = c4_ids = None
    if rank == 0:
        c4_off, c4_ids = W.uniform_label_lists(10_000_000, 65536, 5, dev)
    res["c4"] = run("c4", c4_off, c4_ids, reps=3)
    return res


# ------------------------------------------------------------------ the other configurations (N = 1)

def configs_section(args, ctx, dev, peak):
    """BASELINE.json configs[0..3], the equal-length control and Zipf s = 0.5 of SURVEY 8(d): value, parity against
    the CPU reference (all lists for C1-C4), roofline fraction of the slower ROC kernel. Each leg is independent."""
    import torch

    from vector_db_id_compression_b200 import workloads as W

    res = {}

    def guard(name, fn):
        t0 = time.perf_counter()
        try:
            res[name] = fn()
        except Exception as ex:  # noqa: BLE001
            res[name] = {"error": f"{type(ex).__name__}: {ex}"}
        res[name]["seconds"] = round(time.perf_counter() - t0, 2)
        torch.cuda.empty_cache()

    def ivf(n, nlist, seed, frac, zipf=None, steps=5):
        if zipf is None:
            off, idt = W.uniform_label_lists(n, nlist, seed, dev)
        else:
            off, idt = W.random_partition_lists(n, W.zipf_sizes(n, nlist, zipf), seed, dev)
        t = time_roc(args, ctx, off, idt, dev, steps, 3, 1, torch.cuda.synchronize)
        par, cpu = roc_parity(args, t["blob"], t["out"], off, idt, dev, frac=frac, always_longest=8 if frac < 1 else 0,
                              cpu_seconds=6.0)
        roof = roc_roofline(t["kernel_ms"], n, t["blob"].ans_bytes, peak, "", args, traffic_ok=False)
        bits = 8.0 * t["blob"].ans_bytes / n
        t["blob"].free()
        efr = ef_section(args, ctx, off, idt, dev, peak, cpu=False, steps=3, warmup=2)
        return {"n_ids": n, "nlist": nlist, "value": n * steps / (t["ms_total"] * 1e-3), "unit": "ids/s",
                "ms_per_step": t["ms_total"] / steps, "kernel_ms": {k: round(v, 4) for k, v in t["kernel_ms"].items()},
                "roofline": {"kernel": roof["kernel"], "frac": roof["frac"], "achieved": roof["achieved"]},
                "bits_per_id": bits, "parity": par, "cpu_baseline": {k: cpu[k] for k in ("value", "cores", "kind")},
                "ef": {"encode_frac": efr["encode"]["frac"], "decode_frac": efr["decode"]["frac"],
                       "encode_ms": efr["encode"]["kernel_ms"], "decode_ms": efr["decode"]["kernel_ms"],
                       "roundtrip": efr["bit_exact_roundtrip"]}}

    guard("c1", lambda: config_c1(ctx, dev))
    guard("c2", lambda: ivf(1_000_000, 1024, 2, 1.0))
    guard("c3", lambda: config_c3(args, ctx, dev, peak))
    guard("c4", lambda: ivf(10_000_000, 65536, 5, 1.0))
    n5 = int(args.n_ids)
    guard("control", lambda: ivf(n5, args.nlist, 6, 0.01, zipf=0.0, steps=3))
    guard("s05", lambda: ivf(n5, args.nlist, 7, 0.01, zipf=0.5, steps=3))
    return res


def config_c1(ctx, dev):
    """configs[0]: IVF256,Flat on 100 k vectors of d = 64 (256-byte codes) through the plugin classes (the host-side
    mirror of custom_invlists): per list decoded id set, code <-> id pairing, ROC streams against the CPU reference,
    Elias-Fano in id order (test_compressed_ivfs.py:26-90)."""
    from vector_db_id_compression_b200 import custom_invlists as ci

    codec, kind = cpu_codec()
    rng = np.random.default_rng(1)
    nb, nlist, cs = 100_000, 256, 256
    assign = rng.integers(0, nlist, size=nb)
    codes = rng.integers(0, 256, size=(nb, cs), dtype=np.uint8)
    il = ci.InvertedLists(nlist, cs)
    order = np.argsort(assign, kind="stable")
    bounds = np.searchsorted(assign[order], np.arange(nlist + 1))
    for l in range(nlist):
        idl = order[bounds[l]: bounds[l + 1]]
        il.add_entries(l, idl, codes[idl])
    t0 = time.perf_counter()
    roc = ci.CompressedIDInvertedListsFenwickTree(il, ctx)
    t1 = time.perf_counter()
    ef = ci.CompressedIDInvertedListsEliasFano(il, ctx)
    t2 = time.perf_counter()
    roc.prefetch(range(nlist))
    ef.prefetch(range(nlist))
    t3 = time.perf_counter()
    ex = roc.blob.export()
    off, idc = il.csr()
    prec = ex["precision"].astype(np.uint8)
    heads, nwords, woff, words = codec.encode_lists(off, idc.astype(np.uint64), prec, nthreads=0)
    cdec = codec.decode_lists(off, prec, woff, nwords, heads, words, nthreads=0)
    bad = 0
    for l in range(nlist):
        a, b = int(off[l]), int(off[l + 1])
        got = roc.get_ids(l)
        w0, w1 = int(ex["word_offsets"][l]), int(ex["word_offsets"][l + 1])
        cw = words[int(woff[l]): int(woff[l]) + int(nwords[l])]
        ok = (np.array_equal(np.sort(got), idc[a:b]) and np.array_equal(got.astype(np.uint64), cdec[a:b])
              and int(ex["heads"][l]) == int(heads[l]) and np.array_equal(ex["words"][w0:w1], cw)
              and np.array_equal(roc.get_codes(l), codes[got])            # code j belongs to id j
              and np.array_equal(ef.get_ids(l), idc[a:b]) and np.array_equal(ef.get_codes(l), codes[idc[a:b]]))
        bad += 0 if ok else 1
    return {"n_ids": nb, "nlist": nlist, "code_size": cs, "lists_checked": nlist, "mismatching_lists": bad,
            "bit_exact": bad == 0, "against": kind,
            "what": "plugin classes: ROC (head, words) and decode order vs the reference, id sets, code <-> id pairing, EF in id order",
            "roc_ctor_ms": 1e3 * (t1 - t0), "ef_ctor_ms": 1e3 * (t2 - t1), "decode_all_ms": 1e3 * (t3 - t2),
            "roc_bits_per_id": 8.0 * roc.compressed_ids_size_in_bytes / nb, "ef_bits_per_id": 8.0 * ef.compressed_ids_size_in_bytes / nb}


def config_c3(args, ctx, dev, peak):
    """configs[2]: NSG-like adjacency, 1 M rows x K = 64, int32: row encode, decode of all rows, 10^7 random rows
    (device-resident row numbers), for Elias-Fano and ROC; ROC streams of ALL rows against the CPU reference."""
    import torch

    from vector_db_id_compression_b200 import workloads as W

    N, K = 1_000_000, 64
    data, _ = W.nsg_like_graph(N, K, 3, dev)
    deg = (data >= 0).sum(1)
    edges = int(deg.sum())
    sel = torch.randint(0, N, (10_000_000,), generator=torch.Generator(device=dev).manual_seed(4), device=dev, dtype=torch.int32)
    big = torch.full_like(data, 2**31 - 1)
    srt = torch.sort(torch.where(data >= 0, data, big), dim=1)[0]

    def timed(fn, reps=3):
        best, r = 1e9, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = fn()
            ctx.synchronize()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t)
        return r, best

    out = {"rows": N, "K": K, "edges": edges}
    for name, enc in (("ef", ctx.ef_encode_rows), ("roc", ctx.roc_encode_rows)):
        blob, te = timed(lambda: enc(data), reps=3)
        (nb, cnt), td = timed(lambda: blob.decode_rows(device=dev))
        kms = ctx.last_kernel_ms()
        (nb2, cnt2), tr = timed(lambda: blob.decode_rows(sel, device=dev), reps=2)
        got = torch.sort(torch.where(nb >= 0, nb, big), dim=1)[0]
        ok = bool(torch.equal(got, srt)) and bool(torch.equal(cnt.long(), deg)) and bool(torch.equal(nb2, nb[sel.long()]))
        size = blob.bits_total / 8 if name == "ef" else blob.ans_bytes
        alg = size + 4.0 * edges
        r = {"bits_per_edge": 8 * size / edges, "encode_ms": 1e3 * te, "encode_edges_per_s": edges / te,
             "decode_all_ms": 1e3 * td, "decode_all_kernel_ms": kms, "decode_edges_per_s": edges / td,
             "decode_frac": alg / (kms * 1e-3) / 1e9 / peak if kms else None,
             "random_rows": int(sel.numel()), "random_ms": 1e3 * tr, "random_rows_per_s": sel.numel() / tr, "roundtrip_ok": ok}
        if name == "roc":
            # all rows against the reference: (head, words) per row and the decode order
            codec, kind = cpu_codec()
            ex = blob.export()
            degh = deg.cpu().numpy().astype(np.int64)
            off = np.zeros(N + 1, np.uint64)
            off[1:] = np.cumsum(degh)
            flat = data[data >= 0].cpu().numpy().astype(np.uint64)  # row-major: row i's ids in input order
            prec = ex["precision"].astype(np.uint8)
            heads, nwords, woff, words = codec.encode_lists(off, flat, prec, nthreads=0)
            cdec = codec.decode_lists(off, prec, woff, nwords, heads, words, nthreads=0)
            wo = ex["word_offsets"].astype(np.int64)
            nw = np.asarray(nwords[:N]).astype(np.int64)
            same = bool(np.array_equal(ex["heads"], np.asarray(heads[:N]))) and bool(np.array_equal(np.diff(wo), nw))
            if same:
                rel = np.arange(int(nw.sum()), dtype=np.int64) - np.repeat(np.cumsum(nw) - nw, nw)
                same = bool(np.array_equal(ex["words"], words[np.repeat(np.asarray(woff[:-1]).astype(np.int64), nw) + rel]))
            gflat = nb[nb >= 0].cpu().numpy().astype(np.uint64)
            same = same and bool(np.array_equal(gflat, cdec))
            r["parity"] = {"rows_checked": N, "bit_exact": same, "against": kind,
                           "what": "(head, words) of every row and the decoded order"}
        else:
            r["parity"] = {"rows_checked": N, "value_exact": ok, "what": "decoded rows = sorted input rows (all rows)"}
        out[name] = r
        blob.free()
    return out


# ------------------------------------------------------------------ per-call accessors (N = 1)

def accessors_section(args):
    """The path the reference's drivers time: per-call get_neighbors / get_ids / deferred translation through the C++
    adapter (tools/accessor_bench.cpp, built by build()), with the CPU reference's per-row / per-list decode beside it."""
    exe = ROOT / "tools" / "accessor_bench"
    if not exe.exists():
        return {"error": "tools/accessor_bench not built (run __graft_entry__.build())"}
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"error": f"accessor_bench exit {r.returncode}: {r.stderr[-400:]}"}
    res = json.loads(r.stdout.strip().splitlines()[-1])
    # CPU reference for the same shapes: single-thread per-row decode of K = 64 rows, per-list decode of ~977-id lists
    codec, kind = cpu_codec()
    rng = np.random.default_rng(3)
    for name, n, count, bits in (("row_K64", 64, 20000, 20), ("list_977", 977, 2000, 20)):
        off = np.arange(count + 1, dtype=np.uint64) * n
        ids = np.concatenate([np.sort(rng.choice(1 << bits, size=n, replace=False)) for _ in range(count)]).astype(np.uint64)
        prec = np.full(count, bits, np.uint8)
        heads, nwords, woff, words = codec.encode_lists(off, ids, prec, nthreads=1)
        t0 = time.perf_counter()
        codec.decode_lists(off, prec, woff, nwords, heads, words, nthreads=1)
        dt = time.perf_counter() - t0
        res.setdefault("cpu_reference", {})[name] = {"us_per_call": 1e6 * dt / count, "kind": kind, "threads": 1}
    return res


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
