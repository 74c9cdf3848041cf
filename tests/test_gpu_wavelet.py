"""GPU parity tests of the wavelet-tree flavour (csrc/wt_kernels.cu) through the C ABI against
oracle/wt_oracle.c: every array of the structure word for word, every select, whole-list decode, the plugin
class the way the reference's tests use it (test_compressed_ivfs.py:37-41,128-132), error behaviour.
SDSL is absent, so bit-level parity with sdsl::wt_int is unpinned; select values are exact."""
import numpy as np
import pytest

import oracle
from test_oracle_wt import make_lists

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from vector_db_id_compression_b200.capi import Context

    c = Context(0)
    yield c
    import gc

    gc.collect()  # blobs the tests left to the garbage collector
    try:
        c.close()
    except Exception:  # a blob kept alive by a failed test's traceback: the context refuses to go before it
        pass


def skewed_lists(rng, nlist, n, sort_labels):
    lab = np.minimum(rng.pareto(0.7, size=n).astype(np.int64), nlist - 1)
    if sort_labels:
        lab = np.sort(lab)
    order = np.argsort(lab, kind="stable")
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(np.bincount(lab, minlength=nlist))
    return offsets, order.astype(np.int64), lab.astype(np.uint32)


CASES = [(1, 700, 0), (2, 512, 0), (7, 5000, 0), (256, 100_000, 0), (1000, 70_000, 1), (65, 20_481, 2),
         (1024, 1_000_000, 0), (5000, 1_300_001, 1)]


@pytest.mark.parametrize("path", ["replay", "select"])
@pytest.mark.parametrize("nlist,n,skew", CASES)
def test_wt_structure_and_selects_vs_oracle(ctx, monkeypatch, nlist, n, skew, path):
    monkeypatch.setenv("IDC_WT_DECODE", path)  # whole-list decode: streaming replay of the partitions / select walks
    # S[id] = list_no: through id-range buckets (the path of sequences beyond L2; 2^10-id buckets here) / direct stores
    monkeypatch.setenv("IDC_WT_FILL", "bucket" if path == "replay" else "direct")
    monkeypatch.setenv("IDC_WT_BUCKET_LOG", "10" if n < 500_000 else "14")
    rng = np.random.default_rng(n + nlist)
    if skew == 0:
        offsets, ids, lab = make_lists(rng, nlist, n, empty=(2,) if nlist > 3 else ())
    else:
        offsets, ids, lab = skewed_lists(rng, nlist, n, skew == 2)
    S = oracle.wt.sequence(offsets, ids)
    want = oracle.wt.build(nlist, S)
    blob = ctx.wt_encode(offsets, ids)
    assert (blob.nlist, blob.total_ids, blob.levels) == (nlist, n, want["levels"])
    ex = blob.export()
    assert np.array_equal(ex["list_offsets"], offsets)
    for key in ("bits", "rank", "sel1", "sel0", "start"):
        assert np.array_equal(ex[key], want[key]), key
    assert blob.bits_bytes == want["levels"] * want["nblk"] * 64
    # whole lists, all and a subset, 8- and 4-byte output
    dec, off = blob.decode()
    assert np.array_equal(off, offsets) and np.array_equal(dec, ids)
    pick = rng.permutation(nlist)[: min(nlist, 9)]
    d2, o2 = blob.decode(pick, id_bytes=4)
    for j, l in enumerate(pick):
        assert np.array_equal(d2[int(o2[j]): int(o2[j + 1])], ids[int(offsets[l]): int(offsets[l + 1])])
    # random access = get_single_id: against the input and against the oracle's definition-level select
    nonempty = np.nonzero(np.diff(offsets.astype(np.int64)))[0]
    ql = rng.choice(nonempty, size=2000)
    qo = (rng.random(2000) * np.diff(offsets.astype(np.int64))[ql]).astype(np.int64)
    got = blob.select(ql, qo)
    assert np.array_equal(got, ids[offsets[ql].astype(np.int64) + qo])
    for l, k, g in list(zip(ql, qo, got))[:25]:
        assert oracle.wt.select_seq(S, int(l), int(k)) == int(g)
    # out-of-range queries answer -1
    bad = blob.select([nlist, int(nonempty[0])], [0, int(offsets[nonempty[0] + 1] - offsets[nonempty[0]])])
    assert bad.tolist() == [-1, -1]


def test_wt_device_buffers_and_int32_ids(ctx):
    import torch

    rng = np.random.default_rng(5)
    offsets, ids, _ = make_lists(rng, 300, 200_000)
    blob = ctx.wt_encode(offsets, torch.from_numpy(ids).cuda())
    b32 = ctx.wt_encode(offsets, ids.astype(np.int32))
    e1, e2 = blob.export(), b32.export()
    for key in ("bits", "rank", "sel1", "sel0", "start"):
        assert np.array_equal(e1[key], e2[key])
    dec, _ = blob.decode(device="cuda")
    assert dec.is_cuda and np.array_equal(dec.cpu().numpy(), ids)
    ql = rng.integers(0, 300, size=5000)
    qo = (rng.random(5000) * np.diff(offsets.astype(np.int64))[ql]).astype(np.int64)
    got = blob.select(ql, qo, device="cuda")
    assert np.array_equal(got.cpu().numpy(), ids[offsets[ql].astype(np.int64) + qo])


@pytest.mark.parametrize("fill", ["direct", "bucket"])
def test_wt_rejects_what_the_reference_asserts_and_level_counts(ctx, monkeypatch, fill):
    from vector_db_id_compression_b200.capi import IdcError

    monkeypatch.setenv("IDC_WT_FILL", fill)
    monkeypatch.setenv("IDC_WT_BUCKET_LOG", "4")
    for bad in ([1, 0, 2, 3], [0, 1, 2, 4], [0, 1, 1, 3], [0, 1, -2, 3]):  # order, range, duplicate / hole, negative
        with pytest.raises(IdcError):
            ctx.wt_encode([0, 2, 4], np.array(bad, dtype=np.int64))
    rng = np.random.default_rng(9)
    offsets, ids, _ = make_lists(rng, 50, 40_000)
    for pos, val in ((777, None), (20_000, 40_000), (31_000, "dup")):  # the same inside a larger index
        bad = ids.copy()
        bad[pos] = bad[pos - 1] if val is None else (bad[5] if val == "dup" else val)
        with pytest.raises(IdcError):
            ctx.wt_encode(offsets, bad)
    with pytest.raises(IdcError):
        ctx.wt_encode([0, 2, 4], np.array([0, 3, 1, 2], dtype=np.int64), wt_type=2)  # custom_invlists_impl.cpp:349: 0 or 1
    many = ctx.wt_encode(np.arange(0, 200_001, 2), np.arange(200_000))  # 100 000 lists: 17 levels, 16-bit symbols after level 0
    wide = ctx.wt_encode(np.arange(0, 300_001, 2), np.arange(300_000))  # 150 000 lists: 18 levels, 32-bit symbols
    for bl, m in ((many, 200_000), (wide, 300_000)):
        assert bl.levels == (m // 2 - 1).bit_length()
        assert np.array_equal(bl.decode()[0], np.arange(m))
        assert bl.select([m // 2 - 1, 7], [1, 0]).tolist() == [m - 1, 14]
    blob = ctx.wt_encode([0, 2, 4], np.array([0, 3, 1, 2], dtype=np.int64))  # the context is still usable
    assert blob.decode()[0].tolist() == [0, 3, 1, 2]
    empty = ctx.wt_encode([0, 0, 0], np.zeros(0, np.int64))
    assert empty.total_ids == 0 and empty.decode()[0].size == 0 and empty.select([0], [0]).tolist() == [-1]


def test_plugin_wavelet_tree_like_reference_tests(ctx):
    """test_compressed_ivfs.py:37-41: get_single_id(list, offset) == the list's ids[offset]; :128-132 get_ids."""
    from test_gpu_parity import make_ivf
    from vector_db_id_compression_b200 import custom_invlists as ci

    rng = np.random.default_rng(4)
    il, codes = make_ivf(rng, nlist=8, nb=1000)
    inv = ci.CompressedIDInvertedListsWaveletTree(il, 0, ctx)
    assert inv.nlist == 8 and inv.code_size == 4 and inv.wt_type == 0
    for c in range(8):
        n = inv.list_size(c)
        assert n == il.list_size(c)
        got = inv.get_ids(c)
        assert np.array_equal(got, il.get_ids(c))
        assert np.array_equal(inv.get_codes(c), codes[got])
        for j in (0, n // 2, n - 1):
            assert inv.get_single_id(c, j) == int(il.get_ids(c)[j])
        inv.release_ids(c, got)
    assert 1000 * 3 // 8 <= inv.compressed_ids_size_in_bytes < 8 * 1000  # 3 levels x 1000 bits + directories
    labels = np.array([[(3 << 32) | 5, (0 << 32) | 0, -1], [(7 << 32) | 2, (3 << 32) | 1, (3 << 32) | 5]], dtype=np.int64)
    out = ci.translate_labels(inv, labels)
    for (q, j), lab in np.ndenumerate(labels):
        want = -1 if lab < 0 else int(il.get_ids(int(lab >> 32))[int(lab & 0xFFFFFFFF)])
        assert int(out[q, j]) == want
    assert np.array_equal(ci.translate_labels(inv, labels, decode_1by1=True), out)
    with pytest.raises(Exception):
        ci.CompressedIDInvertedListsWaveletTree(il, 2, ctx)  # custom_invlists_impl.cpp:349: wt_type is 0 or 1


RRR_CASES = [(1, 700, 0), (7, 5000, 0), (256, 100_000, 0), (1000, 70_000, 1), (65, 20_481, 2), (5000, 1_300_001, 1)]


@pytest.mark.parametrize("path", ["replay", "select"])
@pytest.mark.parametrize("nlist,n,skew", RRR_CASES)
def test_wt_type1_rrr_blocks_vs_oracle(ctx, monkeypatch, tmp_path, nlist, n, skew, path):
    """wt_type = 1 (sdsl::wt_int<rrr_vector<63>>, custom_invlists_impl.cpp:371-372): the levels as RRR(63) blocks. The
    compressed arrays word for word against oracle/wt_oracle.c (oracle_rrr_encode), every select / whole-list decode
    against the input and against the plain flavour, the plain levels recovered exactly, the file form."""
    monkeypatch.setenv("IDC_WT_DECODE", path)
    rng = np.random.default_rng(3 * n + nlist)
    if skew == 0:
        offsets, ids, lab = make_lists(rng, nlist, n, empty=(2,) if nlist > 3 else ())
    else:
        offsets, ids, lab = skewed_lists(rng, nlist, n, skew == 2)
    plain = ctx.wt_encode(offsets, ids, wt_type=0)
    blob = ctx.wt_encode(offsets, ids, wt_type=1)
    assert blob.info.wt_type == 1 and (blob.nlist, blob.total_ids, blob.levels) == (nlist, n, plain.levels)
    ex0, ex1 = plain.export(), blob.export()
    for key in ("list_offsets", "bits", "rank", "sel1", "sel0", "start"):  # export gives the plain levels back
        assert np.array_equal(ex0[key], ex1[key]), key
    want = oracle.wt.rrr_encode(ex0["bits"])
    got = blob.export_rrr()
    for key in ("cls", "ptr", "off_base", "off"):
        assert np.array_equal(got[key], want[key]), key
    assert np.array_equal(oracle.wt.rrr_decode(got), ex0["bits"])
    nblk = (n + 511) // 512
    assert blob.bits_bytes == blob.levels * nblk * 8 + blob.levels * (nblk + 1) * 4 + int(want["off_base"][-1]) * 8
    if skew:  # skewed list sizes make skewed upper levels: that is where the block coder saves space
        assert blob.bits_bytes < plain.bits_bytes
    dec, off = blob.decode()
    assert np.array_equal(off, offsets) and np.array_equal(dec, ids)
    some = [l for l in (0, nlist // 2, nlist - 1) if l < nlist]
    d2, o2 = blob.decode(some, id_bytes=4)
    assert np.array_equal(d2.astype(np.int64), np.concatenate([ids[int(offsets[l]): int(offsets[l + 1])] for l in some]))
    q = rng.integers(0, n, size=min(n, 3000))
    ql = lab[ids[q]] if False else np.searchsorted(offsets, q, side="right") - 1  # position q of the CSR -> (list, offset)
    qo = q - offsets[ql].astype(np.int64)
    assert np.array_equal(blob.select(ql, qo), ids[q])
    assert np.array_equal(blob.select(ql, qo), plain.select(ql, qo))
    assert blob.select([nlist], [0]).tolist() == [-1]
    path_f = tmp_path / "index_rrr.wt"
    blob.save(path_f)
    back = ctx.wt_load(path_f)
    assert back.info.wt_type == 1 and back.bits_bytes == blob.bits_bytes
    assert np.array_equal(back.select(ql, qo), ids[q])
    for b in (back, blob, plain):
        b.free()


def test_wt_type1_plugin_class(ctx):
    from vector_db_id_compression_b200 import custom_invlists as ci

    rng = np.random.default_rng(17)
    n, nlist = 3000, 8
    lab = rng.integers(0, nlist, size=n)
    il = ci.InvertedLists(nlist, 4)
    for l in range(nlist):
        sel = np.flatnonzero(lab == l)
        il.add_entries(l, sel.astype(np.int64), rng.integers(0, 255, size=(sel.size, 4), dtype=np.uint8))
    inv = ci.CompressedIDInvertedListsWaveletTree(il, 1, ctx)
    assert inv.wt_type == 1
    for l in range(nlist):
        want = np.flatnonzero(lab == l)
        assert np.array_equal(inv.get_ids(l), want)
        for o in (0, want.size // 2, want.size - 1):
            assert inv.get_single_id(l, o) == want[o]
