"""Generate the committed golden vectors by running the UNMODIFIED reference codec.

Run in the build container (needs /root/reference -> oracle/_ref/libref_roc.so):

    python tests/golden/gen_golden.py

Outputs (committed):
    tests/golden/roc_golden.npz   seeded random sets: ids, precision, head, words,
                                  decoded order, sample order -- all from the reference
    tests/golden/roc_kat.json     the KAT table of SURVEY.md 8(c), re-derived from the
                                  reference here, plus the test_codec.cpp:54-106 workload
                                  (n=65000, 20-bit, seeds 0..9) as head / word count / FNV-1a
    tests/golden/ftree_script.json scripted insert/remove results of the reference
                                  order-statistic tree (fenwick_tree.h) incl. the sequence
                                  of test_fenwick_tree.cpp:16-183
The GPU box has no /root/reference; tests there read only these files.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402

HERE = Path(__file__).resolve().parent


def fnv1a64_words(words) -> int:
    f = 1469598103934665603
    for x in np.asarray(words, dtype=np.uint32).tolist():
        f = ((f ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return f


def mt19937_stream(seed: int, count: int) -> np.ndarray:
    # std::mt19937(seed) == numpy legacy init_genrand seeding
    return np.random.RandomState(seed)._bit_generator.random_raw(count).astype(np.uint64)


def test_codec_ids(seed: int, n: int = 65000, nbits: int = 20) -> np.ndarray:
    """The id stream of test_codec.cpp:62-82: mt() & mask, skipping repeats."""
    raw = mt19937_stream(seed, 4 * n) & ((1 << nbits) - 1)
    _, first = np.unique(raw, return_index=True)
    first.sort()
    ids = raw[first][:n]
    assert ids.size == n
    return ids


def main() -> None:
    ref = oracle.ref
    assert ref is not None, "oracle/_ref/libref_roc.so missing: build it where /root/reference exists"

    rng = np.random.default_rng(20261017)
    cases = []

    def add(ids, p, tag):
        ids = np.asarray(ids, dtype=np.uint64)
        head, words, order = ref.encode(ids, p, want_order=True)
        dec, diag = ref.decode(head, words, ids.size, p, diag=True)
        cases.append(dict(tag=tag, ids=ids, p=p, head=head, words=words, order=order, dec=dec,
                          final_head=diag["final_head"], final_nwords=diag["final_nwords"]))

    # small / edge shapes
    add([0], 0, "single-zero-p0")
    add([1], 0, "single-one-p0-precision-bug")
    add([5], 3, "single")
    add([0, 1], 1, "dense-2")
    add(np.arange(16), 4, "dense-16")
    add(np.arange(256), 8, "dense-256")
    add([7, 8], 3, "max-id-pow2-bug")  # reference rule p=ceil(log2(8))=3 loses the top bit
    add([1024, 3, 77], 10, "max-id-pow2-bug-1024")
    add([4294967295, 17, 2147483648], 32, "p32")
    add([(1 << 40) + 5, 123, 1 << 33], 41, "p41-wide")
    add([3, 3, 9, 9, 9, 1], 4, "duplicates")
    # random sets over the whole precision range
    for t in range(60):
        p = int(rng.integers(1, 33))
        n = int(rng.integers(1, 400))
        n = min(n, 1 << p)
        if (1 << p) <= 4096:
            ids = rng.choice(1 << p, size=n, replace=False)
        else:
            ids = np.unique(rng.integers(0, 1 << p, size=n, dtype=np.uint64))
        add(rng.permutation(ids), p, f"rand-{t}")
    # NSG-row shaped: <= 64 ids below 1e6 (p = 20)
    for t in range(20):
        n = int(rng.integers(16, 65))
        add(rng.choice(1_000_000, size=n, replace=False), 20, f"row-{t}")
    # IVF-list shaped
    add(np.sort(rng.choice(1_000_000, size=977, replace=False)), 20, "ivf1024-list")
    add(np.sort(rng.choice(10_000_000, size=153, replace=False)), 24, "ivf65536-list")
    add(np.sort(rng.choice(1_000_000_000, size=4000, replace=False)), 30, "c5-short")

    flat = {}
    meta = []
    for i, c in enumerate(cases):
        flat[f"ids_{i}"] = c["ids"]
        flat[f"words_{i}"] = c["words"]
        flat[f"order_{i}"] = c["order"]
        flat[f"dec_{i}"] = c["dec"]
        meta.append([c["tag"], c["p"], str(c["head"]), str(c["final_head"]), c["final_nwords"]])
    flat["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(HERE / "roc_golden.npz", **flat)

    # ---- KAT table re-derived from the reference ----
    kat = []
    for ids, p in [
        ([5], 3), ([3, 5], 3), ([0, 1, 2, 3], 2), (list(range(1, 8)), 3), ([100, 200, 300], 9),
        ([999999, 1, 500000, 123456, 777777], 20),
        ([12351235, 49024902, 17781778, 36663666], 26),  # test_codec.cpp:26
    ]:
        head, words = ref.encode(ids, p)
        dec = ref.decode(head, words, len(ids), p)
        kat.append(dict(ids=ids, p=p, head=str(head), stack=words.tolist(), size=8 + 4 * len(words),
                        decoded=dec.tolist()))
    ids = [(i * 7919) % 1000003 for i in range(1, 1001)]
    head, words = ref.encode(ids, 20)
    big = dict(desc="(i*7919) mod 1000003, i=1..1000", p=20, head=str(head), nwords=len(words),
               fnv1a64=str(fnv1a64_words(words)), size=8 + 4 * len(words))
    # ---- test_codec.cpp main(): n=65000 distinct 20-bit ids, seeds 0..9 ----
    tc = []
    for seed in range(10):
        ids = test_codec_ids(seed)
        head, words = ref.encode(ids, 20)
        dec = ref.decode(head, words, ids.size, 20)
        assert set(dec.tolist()) == set(ids.tolist())
        tc.append(dict(seed=seed, n=65000, p=20, head=str(head), nwords=len(words),
                       size=8 + 4 * len(words), fnv1a64=str(fnv1a64_words(words)),
                       dec_fnv1a64=str(fnv1a64_words(dec.astype(np.uint32)))))
        print("test_codec seed", seed, "size", 8 + 4 * len(words))
    json.dump(dict(kat=kat, big=big, test_codec=tc, mt1234=oracle.mt1234(8).tolist()),
              open(HERE / "roc_kat.json", "w"), indent=1)

    # ---- order-statistic tree script (test_fenwick_tree.cpp:16-183 + random) ----
    script = []
    t = oracle.multiset_ref()
    seq = [("i", ord(c)) for c in "babdcecc"] + [("r", k) for k in (6, 1, 3, 4, 0, 1)]
    live = 8 - 6
    r2 = np.random.default_rng(7)
    for _ in range(300):
        if live == 0 or r2.random() < 0.6:
            seq.append(("i", int(r2.integers(0, 40))))
            live += 1
        else:
            seq.append(("r", int(r2.integers(0, live))))
            live -= 1
    for op, arg in seq:
        res = t.insert(arg) if op == "i" else t.remove(arg)
        script.append([op, arg, list(res)])
    json.dump(script, open(HERE / "ftree_script.json", "w"))
    print("wrote", len(cases), "golden cases")


if __name__ == "__main__":
    main()
