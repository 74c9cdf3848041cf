"""TEST INFRASTRUCTURE ONLY.

ctypes front-end for the CPU oracles:

* ``oracle.port``  -- the in-repo C restatement (``roc_oracle.c``, ``ef_oracle.c``,
  ``bits_oracle.c``, ``wt_oracle.c`` -> ``liboracle.so``).
* ``oracle.ref``   -- the UNMODIFIED reference ROC codec compiled in place from
  ``/root/reference`` (``ref_shim.cpp`` + the reference's ``codec.cpp`` ->
  ``_ref/libref_roc.so``); ``None`` when that library has not been built.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package. The product package
(``vector_db_id_compression_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so and (if /root/reference is present) _ref/libref_roc.so."""
    so = _HERE / "liboracle.so"
    stale = so.exists() and any(c.stat().st_mtime > so.stat().st_mtime for c in _HERE.glob("*_oracle.c"))
    if force or stale or not so.exists() or (
        Path("/root/reference/custom_invlist_cpp/codec.cpp").exists()
        and not (_HERE / "_ref" / "libref_roc.so").exists()
    ):
        subprocess.run(["make", "-C", str(_HERE), "-s", "all"], check=True)


def _as_u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).astype(np.uint64, copy=False))


class RocCodec:
    """Uniform access to a ROC CPU implementation (prefix 'oracle_' or 'ref_')."""

    def __init__(self, lib: C.CDLL, kind: str):
        self.lib = lib
        self.kind = kind  # "port" | "reference"
        p = "oracle_" if kind == "port" else "ref_"
        self._p = p
        g = lambda name: getattr(lib, p + name)

        self._max_words = g("roc_max_words")
        self._max_words.restype = C.c_uint64
        self._max_words.argtypes = [C.c_uint64, C.c_int]

        self._prec = g("precision_rule")
        self._prec.restype = C.c_uint64
        self._prec.argtypes = [C.c_uint64]

        if kind == "port":
            self._enc = g("roc_encode")
            self._enc.restype = C.c_int64
            self._enc.argtypes = [C.c_uint64, _u64p, C.c_int, C.POINTER(C.c_uint64), _u32p, C.c_uint64, C.c_void_p]
            self._dec = g("roc_decode")
            self._dec.restype = None
            self._dec.argtypes = [C.c_uint64, _u32p, C.c_uint64, C.c_uint64, C.c_int, _u64p, C.c_void_p]
        else:
            self._enc = g("roc_encode_list")
            self._enc.restype = C.c_int64
            self._enc.argtypes = [
                C.c_uint64, _u64p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), _u32p, C.c_uint64, C.c_void_p,
            ]
            self._compress = g("roc_compress")
            self._compress.restype = C.c_int64
            self._compress.argtypes = [C.c_uint64, _u64p, C.c_int, C.POINTER(C.c_uint64), _u32p, C.c_uint64]
            self._dec = g("roc_decompress")
            self._dec.restype = None
            self._dec.argtypes = [C.c_uint64, _u32p, C.c_uint64, C.c_uint64, C.c_int, _u64p, C.c_void_p, C.c_void_p]

        self._enc_lists = g("roc_encode_lists")
        self._enc_lists.restype = C.c_int
        self._enc_lists.argtypes = [C.c_uint64, _u64p, _u64p, _u8p, _u64p, _u64p, _u32p, _u64p, C.c_int]
        self._dec_lists = g("roc_decode_lists")
        self._dec_lists.restype = None
        self._dec_lists.argtypes = [C.c_uint64, _u64p, _u8p, _u64p, _u64p, _u64p, _u32p, _u64p, C.c_int]

        for name, res, args in [
            ("state_new", C.c_void_p, [C.c_uint64, _u32p, C.c_uint64]),
            ("state_free", None, [C.c_void_p]),
            ("state_head", C.c_uint64, [C.c_void_p]),
            ("state_nwords", C.c_uint64, [C.c_void_p]),
            ("state_words", None, [C.c_void_p, _u32p]),
            ("pop_uniform", C.c_uint64, [C.c_void_p, C.c_uint64]),
            ("push_uniform", None, [C.c_void_p, C.c_uint64, C.c_uint64]),
            ("codec_push", None, [C.c_void_p, C.c_uint64, C.c_int]),
            ("codec_pop", C.c_uint64, [C.c_void_p, C.c_int]),
        ]:
            f = g(name)
            f.restype = res
            f.argtypes = args

    # -- scalar helpers -------------------------------------------------
    def max_words(self, n: int, precision: int) -> int:
        return int(self._max_words(int(n), int(precision)))

    def precision_rule(self, max_id: int) -> int:
        return int(self._prec(int(max_id) & 0xFFFFFFFFFFFFFFFF))

    # -- one list ---------------------------------------------------------
    def encode(self, ids, precision: int, want_order: bool = False, seed: int = 1):
        """-> (head, words[u32]) or (head, words, order[u32])."""
        ids = _as_u64(ids)
        n = ids.size
        cap = self.max_words(n, precision)
        words = np.zeros(cap, dtype=np.uint32)
        head = C.c_uint64(0)
        order = np.zeros(max(n, 1), dtype=np.uint32) if want_order else None
        optr = order.ctypes.data_as(C.c_void_p) if want_order else None
        if self.kind == "port":
            w = self._enc(n, ids, int(precision), C.byref(head), words, cap, optr)
        else:
            w = self._enc(n, ids, int(precision), int(seed), C.byref(head), words, cap, optr)
        if w < 0:
            raise RuntimeError("oracle encode: word buffer too small")
        words = words[:w].copy()
        if want_order:
            return int(head.value), words, order[:n].copy()
        return int(head.value), words

    def compress_in_data_order(self, ids, precision: int):
        """reference ``compress`` verbatim (inserts in data order; O(n^2) on sorted input)."""
        if self.kind == "port":
            return self.encode(ids, precision)
        ids = _as_u64(ids)
        cap = self.max_words(ids.size, precision)
        words = np.zeros(cap, dtype=np.uint32)
        head = C.c_uint64(0)
        w = self._compress(ids.size, ids, int(precision), C.byref(head), words, cap)
        if w < 0:
            raise RuntimeError("oracle compress: word buffer too small")
        return int(head.value), words[:w].copy()

    def decode(self, head: int, words, n: int, precision: int, diag: bool = False):
        words = np.ascontiguousarray(words, dtype=np.uint32)
        wbuf = words if words.size else np.zeros(1, dtype=np.uint32)
        out = np.zeros(max(int(n), 1), dtype=np.uint64)
        if self.kind == "port":
            d = np.zeros(4, dtype=np.uint64)
            self._dec(int(head), wbuf, words.size, int(n), int(precision), out, d.ctypes.data_as(C.c_void_p))
            if diag:
                return out[:n].copy(), dict(
                    final_head=int(d[0]), final_nwords=int(d[1]), draws=int(d[2]), max_rise=int(d[3])
                )
        else:
            fh = C.c_uint64(0)
            fn = C.c_uint64(0)
            self._dec(
                int(head), wbuf, words.size, int(n), int(precision), out,
                C.cast(C.byref(fh), C.c_void_p), C.cast(C.byref(fn), C.c_void_p),
            )
            if diag:
                return out[:n].copy(), dict(final_head=int(fh.value), final_nwords=int(fn.value))
        return out[:n].copy()

    # -- many lists (CSR) -------------------------------------------------
    def encode_lists(self, offsets, ids, precision, nthreads: int = 0):
        """-> (heads[u64], nwords[u64], word_offsets[u64] (capacity slots), words[u32])."""
        offsets = _as_u64(offsets)
        ids = _as_u64(ids)
        precision = np.ascontiguousarray(precision, dtype=np.uint8)
        nlist = offsets.size - 1
        sizes = np.diff(offsets.astype(np.int64))
        caps = (sizes * np.maximum(precision.astype(np.int64), 1) + 31) // 32 + 4
        word_offsets = np.zeros(nlist + 1, dtype=np.uint64)
        np.cumsum(caps, out=word_offsets[1:])
        words = np.zeros(max(int(word_offsets[-1]), 1), dtype=np.uint32)
        heads = np.zeros(max(nlist, 1), dtype=np.uint64)
        nwords = np.zeros(max(nlist, 1), dtype=np.uint64)
        ids_buf = ids if ids.size else np.zeros(1, dtype=np.uint64)
        prec_buf = precision if precision.size else np.zeros(1, dtype=np.uint8)
        rc = self._enc_lists(nlist, offsets, ids_buf, prec_buf, word_offsets, heads, words, nwords, int(nthreads))
        if rc != 0:
            raise RuntimeError("oracle encode_lists: overflow")
        return heads[:nlist], nwords[:nlist], word_offsets, words

    def decode_lists(self, offsets, precision, word_offsets, nwords, heads, words, nthreads: int = 0):
        offsets = _as_u64(offsets)
        precision = np.ascontiguousarray(precision, dtype=np.uint8)
        nlist = offsets.size - 1
        out = np.zeros(max(int(offsets[-1]), 1), dtype=np.uint64)
        pad = lambda a, dt: (np.ascontiguousarray(a, dtype=dt) if np.size(a) else np.zeros(1, dtype=dt))
        self._dec_lists(
            nlist, offsets, pad(precision, np.uint8), _as_u64(word_offsets), pad(nwords, np.uint64),
            pad(heads, np.uint64), pad(words, np.uint32), out, int(nthreads),
        )
        return out[: int(offsets[-1])]

    # -- step-wise state ----------------------------------------------------
    def state(self, head: int = 1 << 31, words=()):
        return _State(self, head, words)


class _State:
    def __init__(self, codec: RocCodec, head: int, words):
        self._c = codec
        g = lambda name: getattr(codec.lib, codec._p + name)
        self._g = g
        w = np.ascontiguousarray(words, dtype=np.uint32)
        wbuf = w if w.size else np.zeros(1, dtype=np.uint32)
        self._h = g("state_new")(int(head), wbuf, w.size)

    def __del__(self):
        if getattr(self, "_h", None):
            self._g("state_free")(self._h)
            self._h = None

    @property
    def head(self) -> int:
        return int(self._g("state_head")(self._h))

    @property
    def words(self) -> np.ndarray:
        n = int(self._g("state_nwords")(self._h))
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._g("state_words")(self._h, out)
        return out[:n].copy()

    def pop_uniform(self, nmax: int) -> int:
        return int(self._g("pop_uniform")(self._h, int(nmax)))

    def push_uniform(self, sym: int, nmax: int) -> None:
        self._g("push_uniform")(self._h, int(sym), int(nmax))

    def codec_push(self, sym: int, precision: int) -> None:
        self._g("codec_push")(self._h, int(sym), int(precision))

    def codec_pop(self, precision: int) -> int:
        return int(self._g("codec_pop")(self._h, int(precision)))


class _MultisetPort:
    """Observable behaviour of the reference order-statistic tree (port)."""

    def __init__(self, lib):
        self._lib = lib
        lib.oracle_mset_new.restype = C.c_void_p
        lib.oracle_mset_free.argtypes = [C.c_void_p]
        lib.oracle_mset_insert.argtypes = [C.c_void_p, C.c_int64, _i64p]
        lib.oracle_mset_remove.argtypes = [C.c_void_p, C.c_int, _i64p]
        lib.oracle_mset_size.argtypes = [C.c_void_p]
        lib.oracle_mset_size.restype = C.c_uint32
        lib.oracle_mset_items.argtypes = [C.c_void_p, _i64p]
        self._h = lib.oracle_mset_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_mset_free(self._h)
            self._h = None

    def insert(self, sym: int):
        o = np.zeros(3, dtype=np.int64)
        self._lib.oracle_mset_insert(self._h, int(sym), o)
        return tuple(int(x) for x in o)  # (symbol, start, freq)

    def remove(self, index: int):
        o = np.zeros(3, dtype=np.int64)
        self._lib.oracle_mset_remove(self._h, int(index), o)
        return tuple(int(x) for x in o)

    def items(self):
        n = int(self._lib.oracle_mset_size(self._h))
        o = np.zeros(max(n, 1), dtype=np.int64)
        self._lib.oracle_mset_items(self._h, o)
        return o[:n].tolist()


class _MultisetRef:
    def __init__(self, lib):
        self._lib = lib
        lib.ref_ftree_new.restype = C.c_void_p
        lib.ref_ftree_free.argtypes = [C.c_void_p]
        lib.ref_ftree_insert.argtypes = [C.c_void_p, C.c_int64, _i64p]
        lib.ref_ftree_remove.argtypes = [C.c_void_p, C.c_int, _i64p]
        self._h = lib.ref_ftree_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_ftree_free(self._h)
            self._h = None

    def insert(self, sym: int):
        o = np.zeros(3, dtype=np.int64)
        self._lib.ref_ftree_insert(self._h, int(sym), o)
        return tuple(int(x) for x in o)

    def remove(self, index: int):
        o = np.zeros(3, dtype=np.int64)
        self._lib.ref_ftree_remove(self._h, int(index), o)
        return tuple(int(x) for x in o)


class EfCodec:
    """Elias-Fano restatement (port only; the reference class cannot be built here)."""

    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.oracle_ef_params.restype = None
        lib.oracle_ef_params.argtypes = [
            C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
        ]
        lib.oracle_ef_encode.restype = C.c_int
        lib.oracle_ef_encode.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u64p, _u64p]
        lib.oracle_ef_decode.restype = None
        lib.oracle_ef_decode.argtypes = [C.c_uint64, C.c_uint32, _u64p, _u64p, _u64p]
        lib.oracle_ef_select.restype = C.c_uint64
        lib.oracle_ef_select.argtypes = [C.c_uint64, C.c_uint32, _u64p, _u64p]

        lib.oracle_ef_encode_lists_mt.restype = C.c_int
        lib.oracle_ef_encode_lists_mt.argtypes = [C.c_uint64, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p, _u64p, C.c_int]
        lib.oracle_ef_decode_lists_mt.restype = None
        lib.oracle_ef_decode_lists_mt.argtypes = [C.c_uint64, _u64p, _u8p, _u64p, _u64p, _u64p, _u64p, _u64p, C.c_int]

    def shapes(self, offsets, ids):
        """Per list of an ascending CSR: universe (= last id, custom_invlists_impl.cpp:262), l, word offsets of the
        two bit vectors (elias_fano.hpp:28-29). Vectorised."""
        offsets = np.asarray(offsets).astype(np.int64)
        ids = _as_u64(ids)
        m = np.diff(offsets)
        nz = m > 0
        uni = np.zeros(m.size, dtype=np.uint64)
        uni[nz] = ids[offsets[1:][nz] - 1]
        q = np.zeros(m.size, dtype=np.uint64)
        q[nz] = uni[nz] // m[nz].astype(np.uint64)
        l = np.zeros(m.size, dtype=np.uint8)
        qq = q.copy()
        for _ in range(64):  # msb(q), 0 for q == 0
            big = qq > 1
            if not big.any():
                break
            l[big] += 1
            qq[big] >>= np.uint64(1)
        low_bits = m.astype(np.uint64) * l.astype(np.uint64)
        high_bits = np.where(nz, m.astype(np.uint64) + 1 + (uni >> l.astype(np.uint64)) + 1, 0).astype(np.uint64)
        low_off = np.zeros(m.size + 1, dtype=np.uint64)
        high_off = np.zeros(m.size + 1, dtype=np.uint64)
        np.cumsum((low_bits + 63) // 64, out=low_off[1:])
        np.cumsum((high_bits + 63) // 64, out=high_off[1:])
        return dict(universe=uni, l=l, low_off=low_off, high_off=high_off, bits_total=int(low_bits.sum() + high_bits.sum()))

    def encode_lists(self, offsets, ids, nthreads: int = 0):
        """All lists of an ascending CSR, OpenMP over lists. -> shapes() + low / high word arrays."""
        sh = self.shapes(offsets, ids)
        off = _as_u64(offsets)
        ids = _as_u64(ids)
        low = np.zeros(max(int(sh["low_off"][-1]), 1), dtype=np.uint64)
        high = np.zeros(max(int(sh["high_off"][-1]), 1), dtype=np.uint64)
        rc = self.lib.oracle_ef_encode_lists_mt(off.size - 1, off, ids if ids.size else np.zeros(1, np.uint64),
                                                sh["universe"] if off.size > 1 else np.zeros(1, np.uint64),
                                                sh["low_off"], sh["high_off"], low, high, int(nthreads))
        if rc != 0:
            raise ValueError("ids must be ascending")
        return dict(sh, low=low, high=high)

    def decode_lists(self, offsets, enc, nthreads: int = 0) -> np.ndarray:
        off = _as_u64(offsets)
        out = np.zeros(max(int(off[-1]), 1), dtype=np.uint64)
        l = enc["l"] if enc["l"].size else np.zeros(1, np.uint8)
        self.lib.oracle_ef_decode_lists_mt(off.size - 1, off, l, enc["low_off"], enc["high_off"], enc["low"], enc["high"],
                                           out, int(nthreads))
        return out[: int(off[-1])]

    def params(self, universe: int, m: int):
        """-> (l, low_bits, high_bits)"""
        l = C.c_uint32(0)
        lb = C.c_uint64(0)
        hb = C.c_uint64(0)
        self.lib.oracle_ef_params(int(universe), int(m), C.byref(l), C.byref(lb), C.byref(hb))
        return int(l.value), int(lb.value), int(hb.value)

    def encode(self, ids_sorted, universe: int | None = None):
        """-> dict(l, m, low_bits, high_bits, low[u64], high[u64])"""
        ids = _as_u64(ids_sorted)
        m = ids.size
        if universe is None:
            universe = int(ids.max()) if m else 0
        l, lb, hb = self.params(universe, m)
        low = np.zeros(max((lb + 63) // 64, 1), dtype=np.uint64)
        high = np.zeros(max((hb + 63) // 64, 1), dtype=np.uint64)
        rc = self.lib.oracle_ef_encode(int(universe), m, ids if m else np.zeros(1, np.uint64), low, high)
        if rc != 0:
            raise ValueError("ids must be ascending and <= universe")
        return dict(
            l=l, m=m, low_bits=lb, high_bits=hb, low=low[: (lb + 63) // 64].copy(), high=high[: (hb + 63) // 64].copy()
        )

    def decode(self, enc) -> np.ndarray:
        out = np.zeros(max(enc["m"], 1), dtype=np.uint64)
        pad = lambda a: a if a.size else np.zeros(1, np.uint64)
        self.lib.oracle_ef_decode(enc["m"], enc["l"], pad(enc["low"]), pad(enc["high"]), out)
        return out[: enc["m"]].copy()

    def select(self, enc, k: int) -> int:
        pad = lambda a: a if a.size else np.zeros(1, np.uint64)
        return int(self.lib.oracle_ef_select(int(k), enc["l"], pad(enc["low"]), pad(enc["high"])))


class BitsCodec:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.oracle_bits_for.restype = C.c_int
        lib.oracle_bits_for.argtypes = [C.c_uint64]
        lib.oracle_bits_pack.restype = None
        lib.oracle_bits_pack.argtypes = [C.c_uint64, _u64p, C.c_int, _u8p, C.c_uint64]
        lib.oracle_bits_unpack.restype = None
        lib.oracle_bits_unpack.argtypes = [C.c_uint64, _u8p, C.c_int, _u64p]
        lib.oracle_bits_get.restype = C.c_uint64
        lib.oracle_bits_get.argtypes = [_u8p, C.c_uint64, C.c_int]

    def bits_for(self, ntotal: int) -> int:
        return int(self.lib.oracle_bits_for(int(ntotal)))

    def pack(self, vals, bits: int, nbytes: int | None = None) -> np.ndarray:
        vals = _as_u64(vals)
        if nbytes is None:
            nbytes = (vals.size * bits + 7) // 8
        out = np.zeros(max(nbytes, 1), dtype=np.uint8)
        self.lib.oracle_bits_pack(vals.size, vals if vals.size else np.zeros(1, np.uint64), bits, out, nbytes)
        return out[:nbytes].copy()

    def unpack(self, code, n: int, bits: int) -> np.ndarray:
        code = np.ascontiguousarray(code, dtype=np.uint8)
        out = np.zeros(max(n, 1), dtype=np.uint64)
        self.lib.oracle_bits_unpack(n, code if code.size else np.zeros(1, np.uint8), bits, out)
        return out[:n].copy()

    def get(self, code, k: int, bits: int) -> int:
        return int(self.lib.oracle_bits_get(np.ascontiguousarray(code, dtype=np.uint8), int(k), bits))


class WtCodec:
    """Wavelet-tree id index restatement (port only; SDSL cannot be built here)."""

    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.oracle_wt_levels.restype = C.c_uint32
        lib.oracle_wt_levels.argtypes = [C.c_uint64]
        lib.oracle_wt_sequence.restype = C.c_int
        lib.oracle_wt_sequence.argtypes = [C.c_uint64, _u64p, _i64p, _u32p]
        lib.oracle_wt_select_seq.restype = C.c_int64
        lib.oracle_wt_select_seq.argtypes = [C.c_uint64, _u32p, C.c_uint32, C.c_uint64]
        lib.oracle_wt_build.restype = None
        lib.oracle_wt_build.argtypes = [C.c_uint64, C.c_uint64, _u32p, _u64p, _u32p, _u32p, _u32p, _u32p]
        lib.oracle_wt_select.restype = C.c_int64
        lib.oracle_wt_select.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, _u32p, C.c_uint32, C.c_uint64]
        lib.oracle_rrr_encode.restype = C.c_uint64
        lib.oracle_rrr_encode.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u64p, _u32p, _u64p, _u64p]
        lib.oracle_rrr_decode.restype = None
        lib.oracle_rrr_decode.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, _u64p, _u64p, _u64p]

    def rrr_encode(self, bits) -> dict:
        """wt_type = 1: the levels' plain bits [levels, nblk * 8] -> dict(cls[levels, nblk], ptr[levels, nblk + 1],
        off_base[levels + 1], off[off_base[-1]]), the RRR(63) block form restated with plain loops."""
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        levels, words = bits.shape
        nblk = words // 8
        cls = np.zeros((levels, max(nblk, 1)), np.uint64)
        ptr = np.zeros((levels, nblk + 1), np.uint32)
        off_base = np.zeros(levels + 1, np.uint64)
        off = np.zeros(levels * (nblk * 8 + 1) + 1, np.uint64)
        used = int(self.lib.oracle_rrr_encode(levels, nblk, bits.reshape(-1) if bits.size else np.zeros(1, np.uint64),
                                              cls.reshape(-1), ptr.reshape(-1), off_base, off))
        return dict(cls=cls[:, :nblk], ptr=ptr, off_base=off_base, off=off[:used])

    def rrr_decode(self, enc: dict) -> np.ndarray:
        cls = np.ascontiguousarray(enc["cls"], dtype=np.uint64)
        levels, nblk = cls.shape
        bits = np.zeros((levels, max(nblk * 8, 1)), np.uint64)
        self.lib.oracle_rrr_decode(levels, nblk, cls.reshape(-1) if cls.size else np.zeros(1, np.uint64),
                                   np.ascontiguousarray(enc["ptr"], dtype=np.uint32).reshape(-1),
                                   np.ascontiguousarray(enc["off_base"], dtype=np.uint64),
                                   np.ascontiguousarray(enc["off"], dtype=np.uint64) if len(enc["off"]) else np.zeros(1, np.uint64),
                                   bits.reshape(-1))
        return bits[:, : nblk * 8]

    def levels(self, nlist: int) -> int:
        return int(self.lib.oracle_wt_levels(int(nlist)))

    def sequence(self, offsets, ids) -> np.ndarray:
        """S[id] = list_no; raises when the lists do not partition [0, ntotal) in ascending order."""
        offsets = _as_u64(offsets)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = int(offsets[-1] - offsets[0])
        S = np.zeros(max(n, 1), dtype=np.uint32)
        if self.lib.oracle_wt_sequence(offsets.size - 1, offsets, ids if ids.size else np.zeros(1, np.int64), S) != 0:
            raise ValueError("lists must be ascending and partition [0, ntotal)")
        return S[:n]

    def select_seq(self, S, c: int, k: int) -> int:
        S = np.ascontiguousarray(S, dtype=np.uint32)
        return int(self.lib.oracle_wt_select_seq(S.size, S if S.size else np.zeros(1, np.uint32), int(c), int(k)))

    def build(self, nlist: int, S):
        """-> dict(levels, n, nblk, bits[levels, words], rank[levels, nblk+1], sel1, sel0, start[nlist])"""
        S = np.ascontiguousarray(S, dtype=np.uint32)
        n = S.size
        levels = self.levels(nlist)
        nblk = (n + 511) // 512
        samp = (n >> 11) + 2
        bits = np.zeros((levels, max(nblk * 8, 1)), dtype=np.uint64)
        rank = np.zeros((levels, nblk + 1), dtype=np.uint32)
        sel1 = np.zeros((levels, samp), dtype=np.uint32)
        sel0 = np.zeros((levels, samp), dtype=np.uint32)
        start = np.zeros(max(nlist, 1), dtype=np.uint32)
        self.lib.oracle_wt_build(int(nlist), n, S if n else np.zeros(1, np.uint32), bits.reshape(-1), rank.reshape(-1),
                                 sel1.reshape(-1), sel0.reshape(-1), start)
        return dict(levels=levels, n=n, nblk=nblk, nlist=int(nlist), bits=bits[:, : nblk * 8], rank=rank, sel1=sel1,
                    sel0=sel0, start=start[:nlist])

    def select(self, wt, c: int, k: int) -> int:
        bits = np.ascontiguousarray(wt["bits"]).reshape(-1)
        return int(self.lib.oracle_wt_select(wt["nlist"], wt["n"], bits if bits.size else np.zeros(1, np.uint64),
                                             np.ascontiguousarray(wt["rank"]).reshape(-1),
                                             np.ascontiguousarray(wt["start"]), int(c), int(k)))


def mt1234(count: int = 8) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint32)
    _port_lib.oracle_mt1234.argtypes = [_u32p, C.c_int]
    _port_lib.oracle_mt1234(out, count)
    return out


build()
_port_lib = C.CDLL(str(_HERE / "liboracle.so"))
port = RocCodec(_port_lib, "port")
ef = EfCodec(_port_lib)
bits = BitsCodec(_port_lib)
wt = WtCodec(_port_lib)
multiset_port = lambda: _MultisetPort(_port_lib)

_ref_path = _HERE / "_ref" / "libref_roc.so"
if _ref_path.exists():
    _ref_lib = C.CDLL(str(_ref_path))
    ref = RocCodec(_ref_lib, "reference")
    multiset_ref = lambda: _MultisetRef(_ref_lib)
    _ref_lib.ref_num_threads.restype = C.c_int
else:  # pragma: no cover - only on a box where the reference never got built
    _ref_lib = None
    ref = None
    multiset_ref = None


def host_threads() -> int:
    return os.cpu_count() or 1
