"""CPU: the Elias-Fano and packed-bits restatements (oracle/ef_oracle.c, bits_oracle.c).

The reference EF class needs ot/succinct (not vendored, cannot be built here), so
these tests pin what the reference's own tests and formulas pin: l = msb(max_id/m)
(elias_fano.hpp:28), the two bit counts (elias_fano.hpp:29,40-42; summed at
custom_invlists_impl.cpp:277), decoded ids == sorted input
(test_compressed_ivfs.py:74-79), select(k) == k-th id (elias_fano.hpp:141-145).
"""
import numpy as np

import oracle


def _py_ef_bits(ids, universe):
    """Independent pure-Python restatement of the builder for small cases."""
    m = len(ids)
    l = (universe // m).bit_length() - 1 if m and universe // m else 0
    low = 0
    high = 0
    for i, v in enumerate(ids):
        low |= (v & ((1 << l) - 1)) << (i * l)
        high |= 1 << ((v >> l) + i)
    return l, low, high, m * l, (m + 1) + (universe >> l) + 1


def test_params_and_bits_small():
    rng = np.random.default_rng(1)
    for trial in range(300):
        m = int(rng.integers(1, 80))
        top = int(rng.integers(m, 1 << int(rng.integers(7, 34))))
        ids = np.sort(rng.choice(top, size=m, replace=False)) if top < 1 << 20 else np.unique(rng.integers(0, top, size=m))
        m = ids.size
        universe = int(ids.max())
        enc = oracle.ef.encode(ids, universe)
        l, low, high, lb, hb = _py_ef_bits([int(x) for x in ids], universe)
        assert (enc["l"], enc["low_bits"], enc["high_bits"]) == (l, lb, hb)
        got_low = sum(int(w) << (64 * k) for k, w in enumerate(enc["low"]))
        got_high = sum(int(w) << (64 * k) for k, w in enumerate(enc["high"]))
        assert got_low == low and got_high == high
        assert np.array_equal(oracle.ef.decode(enc), ids.astype(np.uint64))
        for k in rng.integers(0, m, size=5):
            assert oracle.ef.select(enc, int(k)) == int(ids[k])


def test_shapes_from_survey():
    # C3 row: m=64, max_id ~ 1e6 -> l = 13 ; C5 list: m=15259, max_id ~ 1e9 -> l = 15
    assert oracle.ef.params(999_999, 64)[0] == 13
    assert oracle.ef.params(999_999_999, 15259)[0] == 15
    # m > universe -> l = 0
    assert oracle.ef.params(5, 10)[0] == 0
    l, lb, hb = oracle.ef.params(0, 1)
    assert (l, lb, hb) == (0, 0, 3)


def test_edge_cases():
    for ids in ([0], [0, 1, 2, 3], [7], [5, 5, 5], [0, (1 << 32) + 3], list(range(1000)), [1 << 40]):
        enc = oracle.ef.encode(ids)
        assert oracle.ef.decode(enc).tolist() == ids
        assert [oracle.ef.select(enc, k) for k in range(len(ids))] == ids
    rng = np.random.default_rng(4)
    ids = np.sort(rng.choice(1_000_000_000, size=15259, replace=False))
    enc = oracle.ef.encode(ids)
    assert np.array_equal(oracle.ef.decode(enc), ids.astype(np.uint64))
    # size formula used for compressed_ids_size_in_bytes (custom_invlists_impl.cpp:277)
    assert enc["low_bits"] + enc["high_bits"] == 15259 * 15 + 15259 + 1 + (int(ids.max()) >> 15) + 1


def test_packed_bits():
    b = oracle.bits
    assert [b.bits_for(n) for n in (0, 1, 2, 3, 4, 1000, 1 << 20)] == [0, 1, 2, 2, 3, 10, 21]
    rng = np.random.default_rng(9)
    for bits in (1, 7, 8, 13, 20, 21, 33):
        vals = rng.integers(0, 1 << bits, size=257, dtype=np.uint64)
        code = b.pack(vals, bits)
        assert code.size == (257 * bits + 7) // 8
        assert np.array_equal(b.unpack(code, 257, bits), vals)
        assert b.get(code, 100, bits) == int(vals[100])
    # LSB-first layout: value 0b101 in 3 bits then 0b11 in 3 bits -> 0b011101
    assert b.pack([5, 3], 3).tolist() == [0b011101]
