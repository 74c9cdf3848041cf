"""CPU: the wavelet-tree restatement (oracle/wt_oracle.c).

SDSL is absent (unpinned third-party dependency of the reference), so these tests pin what the reference pins:
get_single_id(list_no, offset) = wt.select(offset + 1, list_no) over S[id] = list_no
(custom_invlists_impl.cpp:354-362,377-379) returns the offset-th id of the list
(test_compressed_ivfs.py:37-41,128-132).
"""
import numpy as np
import pytest

import oracle


def make_lists(rng, nlist, n, empty=()):
    """random partition of [0, n) into nlist ascending lists (CSR)"""
    lab = rng.integers(0, nlist, size=n)
    for e in empty:
        lab[lab == e] = (e + 1) % nlist
    order = np.argsort(lab, kind="stable")
    sizes = np.bincount(lab, minlength=nlist)
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(sizes)
    return offsets, order.astype(np.int64), lab.astype(np.uint32)


def py_wavelet_matrix(S, levels):
    """independent pure-Python restatement for small cases: per level the bit string, then the stable partition"""
    cur = list(int(x) for x in S)
    out = []
    for lev in range(levels):
        sh = levels - 1 - lev
        out.append([(c >> sh) & 1 for c in cur])
        cur = [c for c in cur if not (c >> sh) & 1] + [c for c in cur if (c >> sh) & 1]
    return out, cur


@pytest.mark.parametrize("nlist,n", [(1, 10), (2, 513), (5, 1000), (13, 3000), (64, 5000), (100, 2048), (300, 4097)])
def test_select_is_kth_id(nlist, n):
    rng = np.random.default_rng(nlist * 7 + n)
    offsets, ids, lab = make_lists(rng, nlist, n, empty=(3,) if nlist > 4 else ())
    S = oracle.wt.sequence(offsets, ids)
    assert np.array_equal(S, lab)
    wt = oracle.wt.build(nlist, S)
    levels = wt["levels"]
    assert levels == max(1, int(nlist - 1).bit_length())
    # bits against the pure-Python matrix
    pbits, bottom = py_wavelet_matrix(S, levels)
    for lev in range(levels):
        got = [(int(wt["bits"][lev, i >> 6]) >> (i & 63)) & 1 for i in range(n)]
        assert got == pbits[lev]
        assert int(wt["rank"][lev, -1]) == sum(pbits[lev])
        for j in range(wt["nblk"]):
            assert int(wt["rank"][lev, j]) == sum(pbits[lev][: 512 * j])
    # below the last level every list is one run starting at start[c]
    for c in range(nlist):
        a, b = int(offsets[c]), int(offsets[c + 1])
        s = int(wt["start"][c])
        assert bottom[s : s + (b - a)] == [c] * (b - a)
    # select: definition, wavelet walk, and the input ids agree
    for c in range(nlist):
        a, b = int(offsets[c]), int(offsets[c + 1])
        ks = range(b - a) if b - a <= 40 else rng.integers(0, b - a, size=40)
        for k in ks:
            want = int(ids[a + int(k)])
            assert oracle.wt.select_seq(S, c, int(k)) == want
            assert oracle.wt.select(wt, c, int(k)) == want
        assert oracle.wt.select_seq(S, c, b - a) == -1


def test_rejects_what_the_reference_asserts():
    offsets = np.array([0, 2, 4], np.uint64)
    with pytest.raises(ValueError):
        oracle.wt.sequence(offsets, np.array([1, 0, 2, 3]))  # not ascending (custom_invlists_impl.cpp:358)
    with pytest.raises(ValueError):
        oracle.wt.sequence(offsets, np.array([0, 1, 2, 4]))  # id >= ntotal (:359)
    with pytest.raises(ValueError):
        oracle.wt.sequence(offsets, np.array([0, 1, 1, 3]))  # id owned twice -> another id owned by no list
    assert oracle.wt.sequence(offsets, np.array([0, 3, 1, 2])).tolist() == [0, 1, 1, 0]


def test_property_random_partitions():
    """hypothesis: any partition of [0, n) into ascending lists -> select(c, k) is the k-th id of list c, through the
    wavelet walk and through the definition; access-by-rank consistency of the directories."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 40), st.integers(1, 1500), st.integers(0, 2**31 - 1))
    def check(nlist, n, seed):
        rng = np.random.default_rng(seed)
        # skewed on purpose: most ids in few lists, long runs of equal symbols
        lab = np.minimum((rng.pareto(0.8, size=n)).astype(np.int64), nlist - 1)
        if seed & 1:
            lab = np.sort(lab)
        order = np.argsort(lab, kind="stable").astype(np.int64)
        offsets = np.zeros(nlist + 1, np.uint64)
        offsets[1:] = np.cumsum(np.bincount(lab, minlength=nlist))
        S = oracle.wt.sequence(offsets, order)
        wt = oracle.wt.build(nlist, S)
        # rank directory: last entry of every level = ones of the level; entries ascending, steps <= 512
        for lev in range(wt["levels"]):
            r = wt["rank"][lev].astype(np.int64)
            assert np.all(np.diff(r) >= 0) and np.all(np.diff(r) <= 512)
        for c in rng.integers(0, nlist, size=6):
            a, b = int(offsets[c]), int(offsets[c + 1])
            if b > a:
                k = int(rng.integers(0, b - a))
                assert oracle.wt.select(wt, int(c), k) == int(order[a + k]) == oracle.wt.select_seq(S, int(c), k)

    check()


def test_rrr_blocks_round_trip_and_sizes():
    """wt_type = 1 restatement: plain levels -> RRR(63) blocks -> plain levels, every density; an all-zero or all-one
    block has no offset bits, a balanced one at most 60 (ceil(log2 C(63, 31)))."""
    rng = np.random.default_rng(5)
    for dens in (0.0, 0.02, 0.3, 0.5, 0.9, 1.0):
        bits = np.zeros((4, 5 * 8), np.uint64)
        for l in range(4):
            bits[l] = np.packbits(rng.random(5 * 512) < dens, bitorder="little").view(np.uint64)
        enc = oracle.wt.rrr_encode(bits)
        assert np.array_equal(oracle.wt.rrr_decode(enc), bits)
        per_block = np.diff(enc["ptr"].astype(np.int64), axis=1)
        assert per_block.min() >= 0 and per_block.max() <= 8 * 60
        if dens in (0.0, 1.0):
            assert per_block.max() == 0
        cls = enc["cls"]
        ones = sum(((cls >> np.uint64(6 * j)) & np.uint64(63)).astype(np.int64) for j in range(8))
        tail = np.array([[bin(int(x) >> 48 & 0xFF).count("1") for x in row] for row in cls])
        want = np.array([[bin(int(w)).count("1") for w in row] for row in bits]).reshape(4, 5, 8).sum(axis=2)
        assert np.array_equal(ones + tail, want)
