"""CPU: the device code of the codec steps (csrc/idc_core.cuh, roc_group.cuh, ef_core.cuh), compiled for the host by
tests/hostsim, against the oracle. This checks the exact arithmetic and the order-statistic structures the
CUDA kernels run per lane -- without a GPU. The GPU parity tests proper are tests/test_gpu_parity.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle

HERE = Path(__file__).resolve().parent / "hostsim"
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C")


@pytest.fixture(scope="module")
def sim():
    so = HERE / "libhostsim.so"
    cc = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    subprocess.run([cc, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-o", str(so),
                    str(HERE / "hostsim.cpp")], check=True)
    lib = C.CDLL(str(so))
    enc_args = [C.c_uint32, u64p, C.c_int, C.POINTER(C.c_uint64), u32p, C.c_uint32, u32p, C.POINTER(C.c_uint32)]
    dec_args = [C.c_uint64, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, i64p,
                C.POINTER(C.c_uint32), C.c_uint32]
    lib.sim_group_encode.restype = C.c_int64
    lib.sim_group_encode.argtypes = [C.c_int] + enc_args
    lib.sim_group_decode.restype = None
    lib.sim_group_decode.argtypes = [C.c_int] + dec_args
    lib.sim_warp_encode.restype = C.c_int64
    lib.sim_warp_encode.argtypes = enc_args
    lib.sim_warp_decode.restype = None
    lib.sim_warp_decode.argtypes = [C.c_uint64, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, i64p, C.POINTER(C.c_uint32)]
    lib.sim_small_encode.restype = C.c_int64
    lib.sim_small_encode.argtypes = enc_args
    lib.sim_small_decode.restype = None
    lib.sim_small_decode.argtypes = [C.c_uint64, u32p, C.c_uint32, C.c_uint32, C.c_int, i64p, C.POINTER(C.c_uint32)]
    lib.sim_ef_shape.argtypes = [C.c_uint64, C.c_uint64, u64p]
    lib.sim_ef_encode.argtypes = [i64p, C.c_uint64, C.c_uint64, u64p, u64p, u32p]
    lib.sim_ef_select.restype = C.c_uint64
    lib.sim_ef_select.argtypes = [u64p, u64p, u32p, C.c_uint32, C.c_uint64]
    wt_arrays = [C.c_uint64, C.c_uint64, u64p, u32p, u32p, u32p, u32p]
    lib.sim_wt_levels.restype = C.c_uint32
    lib.sim_wt_levels.argtypes = [C.c_uint64, C.c_uint64]
    lib.sim_wt_build.restype = None
    lib.sim_wt_build.argtypes = [C.c_uint64, C.c_uint64, u32p, u64p, u32p, u32p, u32p, u32p, u64p]
    lib.sim_wt_select.restype = C.c_uint64
    lib.sim_wt_select.argtypes = wt_arrays + [C.c_uint32, C.c_uint64]
    lib.sim_wt_access.restype = C.c_uint32
    lib.sim_wt_access.argtypes = wt_arrays + [C.c_uint64]
    lib.sim_wt_fill_bucketed.restype = C.c_uint32
    lib.sim_wt_fill_bucketed.argtypes = [C.c_uint64, C.c_uint64, i64p, u64p, C.c_uint32, u32p]
    lib.sim_wt_replay_all.restype = None
    lib.sim_wt_replay_all.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, u32p, u64p, i64p]
    lib.sim_wt_replay2_all.restype = None
    lib.sim_wt_replay2_all.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, u32p, u64p, i64p]
    lib.sim_wt_decode_all.restype = None
    lib.sim_wt_decode_all.argtypes = wt_arrays + [u64p, i64p]
    return lib


def grp_enc(lib, G, ids, p):
    ids = np.sort(np.asarray(ids, dtype=np.uint64))
    n = ids.size
    w = np.zeros(n + 4, np.uint32)
    o = np.zeros(max(n, 1), np.uint32)
    h, st = C.c_uint64(), C.c_uint32()
    r = lib.sim_group_encode(G, n, ids, p, C.byref(h), w, n + 4, o, C.byref(st))
    assert r >= 0, st.value
    return h.value, w[:r].copy(), o[:n], st.value


def grp_dec(lib, G, h, w, n, p, lo=0, hi=None, force=0):
    if hi is None:
        hi = (1 << p) - 1 if p < 32 else 0xFFFFFFFF
    out = np.zeros(max(n, 1), np.int64)
    st = C.c_uint32()
    lib.sim_group_decode(G, h, w if w.size else np.zeros(1, np.uint32), w.size, n, p, lo, hi, out, C.byref(st), force)
    return out[:n], st.value


def rand_set(rng, n, p):
    if (1 << p) <= 1 << 16:
        return rng.choice(1 << p, size=min(n, 1 << p), replace=False)
    return np.unique(rng.integers(0, 1 << p, size=n, dtype=np.uint64))


@pytest.mark.parametrize("G", [2, 4, 8])
def test_group_codec_random_sets(sim, G):
    """csrc/roc_group.cuh (G lanes per unit, lanes emulated by host threads) against the oracle."""
    rng = np.random.default_rng(10 + G)
    for trial in range(400):
        p = int(rng.integers(1, 33))
        n = min(int(rng.integers(1, 900)), 1 << p)
        if rng.random() < 0.3:
            p = max(1, int(np.ceil(np.log2(n + 1))))
        ids = rand_set(rng, n, p)
        n = ids.size
        h, w, o = oracle.port.encode(ids, p, want_order=True)
        d = oracle.port.decode(h, w, n, p)
        h2, w2, o2, st = grp_enc(sim, G, ids, p)
        assert (h, w.tolist()) == (h2, w2.tolist()) and st == 0, (trial, n, p)
        srt = np.sort(ids.astype(np.uint64))
        assert np.array_equal(srt[o2], d)
        mode = trial % 3
        if mode == 0:
            d2, st = grp_dec(sim, G, h, w, n, p)
        elif mode == 1:
            d2, st = grp_dec(sim, G, h, w, n, p, lo=int(srt[0]), hi=int(srt[-1]))
        else:
            d2, st = grp_dec(sim, G, h, w, n, p, lo=int(srt[0]), hi=int(srt[-1]), force=1 + int(rng.integers(0, 3)))
        assert np.array_equal(d2.astype(np.uint64), d), (trial, n, p, mode)
        assert st & ~32 == 0


def small_dec(lib, h, w, n, p):
    out = np.zeros(max(n, 1), np.int64)
    st = C.c_uint32()
    lib.sim_small_decode(h, w if w.size else np.zeros(1, np.uint32), w.size, n, p, out, C.byref(st))
    return out[:n], st.value


def test_thread_per_unit_decoder(sim, roc_golden):
    """csrc/roc_small.cuh (one short unit per thread: the graph rows' decoder) against the oracle: random sets of
    1..64 ids (and a few longer ones -- the functions do not depend on the limit) at every precision, sets that fill
    their universe, streams that run into the mt19937 fallback, the golden vectors."""
    rng = np.random.default_rng(64)
    for trial in range(3000):
        p = int(rng.integers(1, 33))
        n = min(int(rng.integers(1, 65 if trial % 10 else 400)), 1 << p)
        if rng.random() < 0.3:
            p = max(1, int(np.ceil(np.log2(n + 1))))
        ids = rand_set(rng, n, p)
        n = ids.size
        h, w = oracle.port.encode(ids, p)
        d2, st = small_dec(sim, h, w, n, p)
        assert np.array_equal(d2.astype(np.uint64), oracle.port.decode(h, w, n, p)) and st == 0, (trial, n, p)
        if n <= 64:  # the encoder twin (a 64-bit presence mask): stream and sample order
            srt = np.sort(ids.astype(np.uint64))
            w2 = np.zeros(n + 4, np.uint32)
            o2 = np.zeros(n, np.uint32)
            h2, st2 = C.c_uint64(), C.c_uint32()
            r = sim.sim_small_encode(n, srt, p, C.byref(h2), w2, n + 4, o2, C.byref(st2))
            assert r == w.size and h2.value == h and np.array_equal(w2[:r], w) and st2.value == 0, (trial, n, p)
            assert np.array_equal(srt[o2], oracle.port.decode(h, w, n, p)), (trial, n, p)
    for c in roc_golden:
        if c["p"] > 32 or c["ids"].size > 5000:
            continue
        d2, st = small_dec(sim, c["head"], c["words"], c["ids"].size, c["p"])
        assert np.array_equal(d2.astype(np.uint64), c["dec"]) and st == 0, c["tag"]
    # precision 0 (max_id = 1: the one id 0) and the reference's power-of-two rule (ids >= 2^p are coded unmasked)
    h, w = oracle.port.encode(np.array([0], np.uint64), 0)
    d2, st = small_dec(sim, h, w, 1, 0)
    assert d2.tolist() == [0] and st == 0


def test_warp_per_unit_coders(sim):
    """The warp-per-unit kernels' bodies (csrc/roc_small.cuh: warp_enc_unit, warp_dec_unit, warp_dec_row; 32 host threads
    play the lanes, the collectives are rendezvous) against the oracle: stream, sample order, decode order; unit lengths
    around the 32-id word and 1024-id half boundaries of the encoder's presence masks, up to the 2 048-id limit."""
    rng = np.random.default_rng(2048)
    sizes = [1, 2, 31, 32, 33, 63, 64, 65, 100, 500, 1023, 1024, 1025, 1500, 2047, 2048] + [int(x) for x in rng.integers(1, 700, size=14)]
    for trial, n in enumerate(sizes):
        p = int(rng.integers(max(1, int(np.ceil(np.log2(n + 1)))), 33))
        ids = rand_set(rng, n, p)
        if ids.size < n:  # (a 64-bit draw may repeat)
            ids = np.unique(np.concatenate([ids, rand_set(rng, n, p)]))[:n]
        n = ids.size
        srt = np.sort(ids.astype(np.uint64))
        h, w = oracle.port.encode(ids, p)
        d = oracle.port.decode(h, w, n, p)
        w2 = np.zeros(n + 4, np.uint32)
        o2 = np.zeros(n, np.uint32)
        h2, st2 = C.c_uint64(), C.c_uint32()
        r = sim.sim_warp_encode(n, srt, p, C.byref(h2), w2, n + 4, o2, C.byref(st2))
        assert r == w.size and h2.value == h and np.array_equal(w2[:r], w) and st2.value == 0, (trial, n, p)
        assert np.array_equal(srt[o2], d), (trial, n, p)
        for row in ((0, 1) if n <= 64 else (0,)):
            out = np.zeros(n, np.int64)
            st = C.c_uint32()
            sim.sim_warp_decode(h, w if w.size else np.zeros(1, np.uint32), w.size, n, p, row, out, C.byref(st))
            assert np.array_equal(out.astype(np.uint64), d) and st.value == 0, (trial, n, p, row)


@pytest.mark.parametrize("G,n,p", [(4, 15259, 30), (4, 65536, 17), (8, 65536, 31), (4, 65000, 20), (8, 4097, 13), (4, 2233, 12), (2, 65536, 30), (2, 15259, 20)])
def test_group_codec_large_units(sim, G, n, p):
    rng = np.random.default_rng(n + p)
    ids = rng.choice(1 << p, size=n, replace=False) if p <= 24 else rand_set(rng, n, p)
    n = ids.size
    h, w = oracle.port.encode(ids, p)
    h2, w2, _, st = grp_enc(sim, G, ids, p)
    assert (h, st) == (h2, 0) and np.array_equal(w, w2)
    d2, st = grp_dec(sim, G, h, w, n, p)
    assert np.array_equal(d2.astype(np.uint64), oracle.port.decode(h, w, n, p)) and st == 0


@pytest.mark.parametrize("G", [2, 4, 8])
def test_group_codec_golden_and_adversarial(sim, roc_golden, G):
    sim_enc = lambda lib, ids, p: grp_enc(lib, G, ids, p)
    sim_dec = lambda lib, h, w, n, p, **kw: grp_dec(lib, G, h, w, n, p, **kw)
    for c in roc_golden:
        if c["p"] > 32:
            continue
        h2, w2, _, _ = sim_enc(sim, c["ids"], c["p"])
        assert h2 == c["head"] and np.array_equal(w2, c["words"]), c["tag"]
        d2, _ = sim_dec(sim, c["head"], c["words"], c["ids"].size, c["p"])
        assert np.array_equal(d2.astype(np.uint64), c["dec"]), c["tag"]
    ids = np.arange(5000, dtype=np.uint64) + (1 << 29)  # one narrow cluster, no range hint: degenerate path
    h, w = oracle.port.encode(ids, 30)
    d = oracle.port.decode(h, w, 5000, 30)
    d2, st = sim_dec(sim, h, w, 5000, 30)
    assert np.array_equal(d2.astype(np.uint64), d) and st == 32
    d2, st = sim_dec(sim, h, w, 5000, 30, lo=1 << 29, hi=(1 << 29) + 4999)
    assert np.array_equal(d2.astype(np.uint64), d) and st == 0


def test_ef_gather_words(sim):
    rng = np.random.default_rng(1)
    for trial in range(250):
        m = int(rng.integers(1, 3000))
        top = max(m + 1, int(rng.integers(1, 1 << int(rng.integers(8, 33)))))
        ids = np.sort(rng.choice(top, size=m, replace=False)) if top < 1 << 22 else np.unique(rng.integers(0, top, size=m))
        m, uni = ids.size, int(ids.max())
        enc = oracle.ef.encode(ids, uni)
        sh = np.zeros(6, np.uint64)
        sim.sim_ef_shape(uni, m, sh)
        assert (int(sh[0]), int(sh[1]), int(sh[2])) == (enc["l"], enc["low_bits"], enc["high_bits"])
        low = np.zeros(max(int(sh[3]), 1), np.uint64)
        high = np.zeros(max(int(sh[4]), 1), np.uint64)
        smp = np.zeros(max(int(sh[5]), 1), np.uint32)
        sim.sim_ef_encode(ids.astype(np.int64), m, uni, low, high, smp)
        assert np.array_equal(low[: int(sh[3])], enc["low"]) and np.array_equal(high[: int(sh[4])], enc["high"])
        for k in rng.integers(0, m, size=8):
            assert sim.sim_ef_select(low, high, smp, enc["l"], int(k)) == int(ids[k])


@pytest.mark.parametrize("nlist,n,skew", [(1, 700, 0), (2, 512, 0), (7, 5000, 0), (256, 100_000, 0), (1000, 70_000, 1),
                                          (65, 20_481, 2)])
def test_wavelet_matrix_build_and_select(sim, nlist, n, skew):
    """csrc/wt_core.cuh + the lane-level build passes of wt_kernels.cu against oracle/wt_oracle.c: every array of
    the structure word for word, then wt_select / wt_access for every id."""
    from test_oracle_wt import make_lists

    rng = np.random.default_rng(n + nlist)
    if skew == 0:
        offsets, ids, lab = make_lists(rng, nlist, n, empty=(2,) if nlist > 3 else ())
    else:
        # skewed list sizes (long runs of one bit value: blocks without ones / zeros, sparse samples)
        lab = np.minimum((rng.pareto(0.7, size=n)).astype(np.int64), nlist - 1)
        if skew == 2:
            lab = np.sort(lab)  # ids of a list are consecutive: whole blocks of equal bits
        order = np.argsort(lab, kind="stable")
        offsets = np.zeros(nlist + 1, np.uint64)
        offsets[1:] = np.cumsum(np.bincount(lab, minlength=nlist))
        ids, lab = order.astype(np.int64), lab.astype(np.uint32)
    S = oracle.wt.sequence(offsets, ids)
    want = oracle.wt.build(nlist, S)
    levels, nblk, samp = want["levels"], want["nblk"], (n >> 11) + 2
    assert sim.sim_wt_levels(nlist, n) == levels
    bits = np.zeros(levels * nblk * 8, np.uint64)
    rank = np.zeros(levels * (nblk + 1), np.uint32)
    sel1 = np.zeros(levels * samp, np.uint32)
    sel0 = np.zeros(levels * samp, np.uint32)
    start = np.zeros(nlist, np.uint32)
    sizes = np.diff(offsets.astype(np.int64)).astype(np.uint64)
    sim.sim_wt_build(nlist, n, np.ascontiguousarray(S), bits, rank, sel1, sel0, start, np.ascontiguousarray(sizes))
    assert np.array_equal(bits.reshape(levels, -1), want["bits"])
    assert np.array_equal(rank.reshape(levels, -1), want["rank"])
    assert np.array_equal(sel1.reshape(levels, -1), want["sel1"])
    assert np.array_equal(sel0.reshape(levels, -1), want["sel0"])
    assert np.array_equal(start, want["start"])
    out = np.zeros(n, np.int64)
    sim.sim_wt_decode_all(nlist, n, bits, rank, sel1, sel0, start, offsets, out)
    assert np.array_equal(out, ids)
    out2 = np.zeros(n, np.int64)
    sim.sim_wt_replay_all(nlist, n, bits, rank, start, offsets, out2)
    assert np.array_equal(out2, ids)
    out3 = np.zeros(n, np.int64)  # two levels per pass, the index arithmetic of k_wt_replay2
    sim.sim_wt_replay2_all(nlist, n, bits, rank, start, offsets, out3)
    assert np.array_equal(out3, ids)
    for i in rng.integers(0, n, size=300):
        assert sim.sim_wt_access(nlist, n, bits, rank, sel1, sel0, start, int(i)) == int(S[i])
    for c, k in [(0, 0), (nlist - 1, 0)]:
        if offsets[c + 1] > offsets[c]:
            assert sim.sim_wt_select(nlist, n, bits, rank, sel1, sel0, start, c, k) == int(ids[int(offsets[c]) + k])


def test_wavelet_bucketed_fill(sim):
    """k_wt_distribute / k_wt_apply, CTA by CTA: S[id] = list_no through id-range buckets equals the direct
    definition (oracle.wt.sequence); the reference's asserts and holes / duplicates come back as status bits."""
    from test_oracle_wt import make_lists

    rng = np.random.default_rng(77)
    for nlist, n, blog in [(50, 40_000, 5), (7, 10_000, 10), (300, 70_001, 9), (3, 5000, 23)]:
        offsets, ids, _ = make_lists(rng, nlist, n)
        S = np.zeros(n, np.uint32)
        assert sim.sim_wt_fill_bucketed(nlist, n, ids, offsets, blog, S) == 0
        assert np.array_equal(S, oracle.wt.sequence(offsets, ids))
        for pos, val, bit in ((777, "prev", 4), (2000, n, 2), (3100, -5, 2), (4000, "dup", 1)):
            bad = ids.copy()
            bad[pos] = bad[pos - 1] if val == "prev" else (bad[5] if val == "dup" else val)
            st = sim.sim_wt_fill_bucketed(nlist, n, bad, offsets, blog, S)
            assert st & bit or st & 4, (nlist, n, pos, val, st)  # a duplicate may also break the order
            assert st != 0
