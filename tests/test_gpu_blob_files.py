"""Blob import and the flat file form (include/idcodec.h: idc_*_blob_save / _load, idc_ef_blob_import,
idc_wt_blob_import): a blob that went through export -> import, or through a file, is indistinguishable from the
original -- same exported arrays, same decoded ids, same random access."""
import numpy as np
import pytest

from test_gpu_parity import make_lists

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


def same_dict(a, b, keys):
    for k in keys:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k


def test_roc_file_round_trip(ctx, tmp_path):
    rng = np.random.default_rng(5)
    offsets, ids = make_lists(rng, [300, 0, 70_000, 1, 4097, 0], 27)  # a multi-unit list, empty lists, n = 1
    blob = ctx.roc_encode(offsets, ids, sorted_ids=True)
    path = tmp_path / "index.roc"
    blob.save(path)
    back = ctx.roc_load(path)
    same_dict(blob.export(), back.export(), ("list_offsets", "unit_offsets", "unit_n", "precision", "heads", "word_offsets", "words"))
    assert np.array_equal(back.decode()[0], blob.decode()[0])
    assert np.array_equal(back.decode([4, 0])[0], blob.decode([4, 0])[0])
    back.free()
    blob.free()


def test_roc_row_blob_file_round_trip(ctx, tmp_path):
    rng = np.random.default_rng(6)
    n, K = 500, 32
    data = np.full((n, K), -1, np.int32)
    for r in range(n):
        c = int(rng.integers(0, K + 1))
        data[r, :c] = rng.choice(n, size=c, replace=False)
    blob = ctx.roc_encode_rows(data)
    path = tmp_path / "graph.roc"
    blob.save(path)
    back = ctx.roc_load(path)
    assert back.row_stride == K
    rows = np.array([0, 17, 499, 3], np.int32)
    a, ca = blob.decode_rows(rows)
    b, cb = back.decode_rows(rows)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)
    assert np.array_equal(back.decode_rows()[0], blob.decode_rows()[0])
    back.free()
    blob.free()


@pytest.mark.parametrize("device_arrays", [False, True])
def test_ef_import_and_file_round_trip(ctx, tmp_path, device_arrays):
    import torch

    rng = np.random.default_rng(7)
    offsets, ids = make_lists(rng, [1, 0, 255, 256, 257, 1024, 40_000, 3, 300_000], 30)
    blob = ctx.ef_encode(offsets, ids, sorted_ids=True)
    ex = blob.export()
    low, high = ex["low"], ex["high"]
    if device_arrays:
        low, high = torch.from_numpy(low.view(np.int64)).cuda(), torch.from_numpy(high.view(np.int64)).cuda()
    imp = ctx.ef_import(ex["list_offsets"], ex["universe"], low, high)
    same_dict(ex, imp.export(), ("list_offsets", "l", "universe", "low_offsets", "high_offsets", "low", "high"))
    assert np.array_equal(imp.decode()[0], ids)
    assert np.array_equal(imp.decode([8, 5])[0], blob.decode([8, 5])[0])
    # select goes through the rebuilt samples: every 255-th / 256-th / 257-th offset of the long lists
    ln, off = [], []
    for l in (6, 8, 2, 3, 4):
        m = int(offsets[l + 1] - offsets[l])
        for o in sorted(set(list(range(0, m, 255)) + list(range(0, m, 256)) + list(range(0, m, 257)) + [m - 1])):
            ln.append(l)
            off.append(o)
    want = np.array([ids[int(offsets[l]) + o] for l, o in zip(ln, off)], np.int64)
    assert np.array_equal(imp.select(ln, off), want)
    path = tmp_path / "index.ef"
    imp.save(path)
    back = ctx.ef_load(path)
    same_dict(ex, back.export(), ("list_offsets", "l", "universe", "low", "high"))
    assert np.array_equal(back.decode()[0], ids)
    assert np.array_equal(back.select(ln, off), want)
    for b in (back, imp, blob):
        b.free()


def test_ef_row_blob_file_round_trip(ctx, tmp_path):
    rng = np.random.default_rng(8)
    n, K = 700, 64
    data = np.full((n, K), -1, np.int32)
    for r in range(n):
        c = int(rng.integers(0, K + 1))
        data[r, :c] = rng.choice(n, size=c, replace=False)
    blob = ctx.ef_encode_rows(data)
    path = tmp_path / "graph.ef"
    blob.save(path)
    back = ctx.ef_load(path)
    assert back.row_stride == K
    rows = np.array([5, 699, 0, 123], np.int32)
    a, ca = blob.decode_rows(rows)
    b, cb = back.decode_rows(rows)
    assert np.array_equal(a, b) and np.array_equal(ca, cb)
    back.free()
    blob.free()


def test_ef_import_refuses_inconsistent_bits(ctx):
    rng = np.random.default_rng(9)
    offsets, ids = make_lists(rng, [500, 20], 20)
    blob = ctx.ef_encode(offsets, ids, sorted_ids=True)
    ex = blob.export()
    high = ex["high"].copy()
    high[0] &= ~np.uint64(int(high[0]) & -int(high[0]))  # drop one set bit
    with pytest.raises(Exception):
        ctx.ef_import(ex["list_offsets"], ex["universe"], ex["low"], high)
    blob.free()


def test_wt_import_and_file_round_trip(ctx, tmp_path):
    rng = np.random.default_rng(10)
    nlist, n = 37, 50_000
    lab = rng.integers(0, nlist, size=n)
    ids = np.argsort(lab, kind="stable").astype(np.int64)
    offsets = np.zeros(nlist + 1, np.uint64)
    offsets[1:] = np.cumsum(np.bincount(lab, minlength=nlist))
    blob = ctx.wt_encode(offsets, ids)
    ex = blob.export()
    imp = ctx.wt_import(ex)
    path = tmp_path / "index.wt"
    imp.save(path)
    back = ctx.wt_load(path)
    for other in (imp, back):
        same_dict(ex, other.export(), ("list_offsets", "bits", "rank", "sel1", "sel0", "start"))
        assert np.array_equal(other.decode()[0], ids)
        q_l = rng.integers(0, nlist, size=200)
        q_o = np.array([rng.integers(0, max(1, int(offsets[l + 1] - offsets[l]))) for l in q_l])
        assert np.array_equal(other.select(q_l, q_o), blob.select(q_l, q_o))
    for b in (back, imp, blob):
        b.free()


def test_load_refuses_wrong_files(ctx, tmp_path):
    rng = np.random.default_rng(11)
    offsets, ids = make_lists(rng, [100, 50], 20)
    blob = ctx.roc_encode(offsets, ids, sorted_ids=True)
    path = tmp_path / "x.roc"
    blob.save(path)
    with pytest.raises(Exception):
        ctx.ef_load(path)  # a ROC file is not an Elias-Fano file
    raw = path.read_bytes()
    (tmp_path / "cut.roc").write_bytes(raw[: len(raw) // 2])
    with pytest.raises(Exception):
        ctx.roc_load(tmp_path / "cut.roc")
    (tmp_path / "junk.roc").write_bytes(b"not a blob at all")
    with pytest.raises(Exception):
        ctx.roc_load(tmp_path / "junk.roc")
    blob.free()
