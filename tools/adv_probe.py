"""Decode of the adversarial 'geo' list (overflowing buckets, spill list, brute-force fallback) against the oracle:
first mismatching step. usage: python tools/adv_probe.py"""
import sys, numpy as np
sys.path.insert(0, "/root/repo")
import oracle
from vector_db_id_compression_b200.capi import Context
ctx = Context(0)
geo = np.unique((1.0003 ** np.arange(1, 60000)).astype(np.int64))[:20000]
for name, ids in (("geo", geo), ("geo8k", geo[:8000]), ("geo3k", geo[:3000])):
    off = np.array([0, ids.size], np.uint64)
    blob = ctx.roc_encode(off, ids, sorted_ids=True)
    p = oracle.port.precision_rule(int(ids[-1]))
    head, words = oracle.port.encode(ids.astype(np.uint64), p)
    want = oracle.port.decode(head, words, ids.size, p)
    got = blob.decode()[0].astype(np.uint64)
    bad = np.nonzero(got != want)[0]
    if bad.size:
        k = int(bad.max())
        print("   first wrong step: got", int(got[k]), "want", int(want[k]), "diff", int(got[k]) - int(want[k]), "| step before: id", int(want[k + 1]),
              "true rank", int((want[k + 2:] < want[k + 1]).sum()))
    print(name, "n", ids.size, "mismatches", bad.size, "last bad index", int(bad.max()) if bad.size else None,
          "= step", ids.size - 1 - int(bad.max()) if bad.size else None)

