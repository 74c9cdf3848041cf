"""Host-side mirror of the reference's `custom_invlists` module (custom_invlist_cpp/custom_invlists.swig,
custom_invlists_impl.{h,cpp}): same class names, same public attributes, same call sequence -- with the
encode / decode loops replaced by bulk calls into the sm_100a codec through the C ABI.

Faiss is not available in this image, so `InvertedLists` below is the minimal stand-in for
faiss::ArrayInvertedLists (nlist, code_size, list_size / get_ids / get_codes); with a real Faiss the same
calls are made from csrc/plugin/idc_faiss_plugin.h.

Differences from the reference, all deliberate and documented in DESIGN.md:
  * the per-list virtuals sit on top of a bulk decode + cache (`prefetch`), because a kernel launch per
    get_ids call cannot win; get_ids(list_no) without a prefetch decodes just that list.
  * lists longer than 65 536 ids are stored as several ROC units (the reference cannot round-trip them).
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np

from . import capi

_default_ctx: Optional[capi.Context] = None


def default_context() -> capi.Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = capi.Context(0)
    return _default_ctx


class InvertedLists:
    """Minimal faiss::ArrayInvertedLists stand-in: per list an int64 id array and a uint8 code array."""

    def __init__(self, nlist: int, code_size: int):
        self.nlist = int(nlist)
        self.code_size = int(code_size)
        self.ids = [np.zeros(0, np.int64) for _ in range(nlist)]
        self.codes = [np.zeros((0, code_size), np.uint8) for _ in range(nlist)]

    def add_entries(self, list_no: int, ids, codes) -> None:
        ids = np.asarray(ids, dtype=np.int64)
        codes = np.asarray(codes, dtype=np.uint8).reshape(ids.size, self.code_size)
        self.ids[list_no] = np.concatenate([self.ids[list_no], ids])
        self.codes[list_no] = np.concatenate([self.codes[list_no], codes])

    def list_size(self, list_no: int) -> int:
        return int(self.ids[list_no].size)

    def get_ids(self, list_no: int) -> np.ndarray:
        return self.ids[list_no]

    def get_codes(self, list_no: int) -> np.ndarray:
        return self.codes[list_no]

    def compute_ntotal(self) -> int:
        return int(sum(x.size for x in self.ids))

    def csr(self):
        offsets = np.zeros(self.nlist + 1, np.uint64)
        offsets[1:] = np.cumsum([x.size for x in self.ids])
        ids = np.concatenate(self.ids) if self.nlist else np.zeros(0, np.int64)
        return offsets, np.ascontiguousarray(ids, dtype=np.int64)


class InvertedListsArrayCodes:
    """custom_invlists_impl.h:22-33: every flavour keeps the codes in plain arrays."""

    def __init__(self, il: InvertedLists):
        self.nlist = il.nlist
        self.code_size = il.code_size
        self.codes_all: list[np.ndarray] = [np.zeros((0, il.code_size), np.uint8)] * il.nlist
        self._cache: dict[int, np.ndarray] = {}

    def list_size(self, list_no: int) -> int:  # custom_invlists_impl.cpp:22-24
        return int(self.codes_all[list_no].shape[0])

    def get_codes(self, list_no: int) -> np.ndarray:  # :26-28
        return self.codes_all[list_no]

    def release_ids(self, list_no: int, ids) -> None:  # delete[] in the reference
        return None

    def get_single_code(self, list_no: int, offset: int) -> np.ndarray:
        return self.codes_all[list_no][offset]

    # bulk decode + cache: the only efficient way to serve many get_ids calls from a GPU codec
    def prefetch(self, list_nos: Iterable[int]) -> None:
        want = sorted({int(l) for l in list_nos if int(l) not in self._cache and self.list_size(int(l))})
        if not want:
            return
        ids, off = self._decode_lists(want)
        for j, l in enumerate(want):
            self._cache[l] = ids[int(off[j]): int(off[j + 1])]

    def drop_cache(self) -> None:
        self._cache.clear()

    def get_ids(self, list_no: int):
        if self.list_size(list_no) == 0:
            return None  # custom_invlists_impl.cpp:212-214, :294-296
        if list_no not in self._cache:
            ids, _ = self._decode_lists([list_no])
            return ids
        return self._cache[list_no]

    def _decode_lists(self, list_nos: Sequence[int]):
        raise NotImplementedError


class CompressedIDInvertedListsFenwickTree(InvertedListsArrayCodes):
    """ROC-compressed ids (custom_invlists_impl.h:56-70, .cpp:133-223)."""

    def __init__(self, il: InvertedLists, ctx: Optional[capi.Context] = None, precision_safe: bool = False):
        super().__init__(il)
        self.ctx = ctx or default_context()
        offsets, ids = il.csr()
        is_sorted = all(bool(np.all(np.diff(x) >= 0)) for x in il.ids if x.size > 1)
        self.blob = self.ctx.roc_encode(offsets, ids, sorted_ids=is_sorted, want_order=True,
                                        precision_safe=precision_safe)
        order = self.blob.order()
        # codes follow the sample order (custom_invlists_impl.cpp:189-193): decoded ids[t] pairs with codes[t]
        self.codes_all = []
        for l in range(il.nlist):
            s, e = int(offsets[l]), int(offsets[l + 1])
            self.codes_all.append(np.ascontiguousarray(il.codes[l][order[s:e]]))
        self._offsets = offsets
        ex = self.blob.export()
        self._ex = ex
        uo = ex["unit_offsets"]
        self.id_symbol_precision = [int(ex["precision"][int(uo[l])]) if e > s else 0
                                    for l, (s, e) in enumerate(zip(offsets[:-1], offsets[1:]))]
        self.compressed_ids_size_in_bytes = self.blob.ans_bytes  # sum of ANSState::size(), :199-202
        self.codes_size_in_bytes = int(sum(c.size for c in self.codes_all))
        self.overhead_in_bytes = 0  # never written by the reference's IVF classes (custom_invlists_impl.h:63)

    @property
    def ans_states(self):
        """[(head, stack words)] per list (first unit), as ANSState{head, stack} (codec.h:13-15)."""
        ex, out = self._ex, []
        for l in range(self.nlist):
            u = int(ex["unit_offsets"][l])
            w0, w1 = int(ex["word_offsets"][u]), int(ex["word_offsets"][u + 1])
            out.append((int(ex["heads"][u]), ex["words"][w0:w1]))
        return out

    def _decode_lists(self, list_nos):
        return self.blob.decode(list_nos)


class CompressedIDInvertedListsEliasFano(InvertedListsArrayCodes):
    """Elias-Fano ids (custom_invlists_impl.h:72-98, .cpp:229-339): ids and codes re-laid in id order."""

    def __init__(self, il: InvertedLists, ctx: Optional[capi.Context] = None):
        super().__init__(il)
        self.ctx = ctx or default_context()
        offsets, ids = il.csr()
        self.codes_all = []
        sorted_ids = []
        for l in range(il.nlist):  # canonicalize_order_inplace (:324-339): sort (id, code) pairs by id
            perm = np.argsort(il.ids[l], kind="stable")
            sorted_ids.append(il.ids[l][perm])
            self.codes_all.append(np.ascontiguousarray(il.codes[l][perm]))
        ids = np.concatenate(sorted_ids) if il.nlist else ids
        self.blob = self.ctx.ef_encode(offsets, np.ascontiguousarray(ids, dtype=np.int64), sorted_ids=True)
        self.compressed_ids_size_in_bytes = self.blob.bits_total // 8  # :277,282
        self.codes_size_in_bytes = int(sum(c.size for c in self.codes_all))
        self.overhead_in_bytes = 0

    def _decode_lists(self, list_nos):
        return self.blob.decode(list_nos)

    def get_single_id(self, list_no: int, offset: int) -> int:  # :314-318
        return int(self.blob.select([list_no], [offset])[0])

    def get_single_ids(self, list_nos, offsets) -> np.ndarray:
        return self.blob.select(list_nos, offsets)


class CompressedIDInvertedListsPackedBits(InvertedListsArrayCodes):
    """Fixed-width ids (custom_invlists_impl.h:37-53, .cpp:62-118)."""

    def __init__(self, il: InvertedLists, ctx: Optional[capi.Context] = None):
        super().__init__(il)
        self.ctx = ctx or default_context()
        ntotal = il.compute_ntotal()
        self.bits = 0
        while (1 << self.bits) < ntotal + 1:  # :66-68
            self.bits += 1
        self.ids_all = []
        self.codes_all = []
        self.compressed_ids_size_in_bytes = 0
        for l in range(il.nlist):
            ids = il.ids[l]
            if ids.size and (ids.min() < 0 or ids.max() >= ntotal):
                raise RuntimeError("Error: 'ids_in[i] >= 0 && ids_in[i] < ntotal' failed")  # FAISS_THROW_IF_NOT :87
            nbytes = (ids.size * self.bits + 7) // 8
            self.ids_all.append(self.ctx.bits_pack(ids.astype(np.uint64), max(self.bits, 1), nbytes) if ids.size
                                else np.zeros(0, np.uint8))
            self.compressed_ids_size_in_bytes += nbytes
            self.codes_all.append(il.codes[l].copy())

    def _decode_lists(self, list_nos):
        parts = [self.ctx.bits_unpack(self.ids_all[l], self.list_size(l), max(self.bits, 1)).astype(np.int64)
                 for l in list_nos]
        off = np.zeros(len(parts) + 1, np.uint64)
        off[1:] = np.cumsum([p.size for p in parts])
        return (np.concatenate(parts) if parts else np.zeros(0, np.int64)), off

    def get_single_id(self, list_no: int, offset: int) -> int:  # BitstringReader_get_bits, :35-58
        code = self.ids_all[list_no]
        pos = offset * self.bits
        v = 0
        for b in range(self.bits):
            v |= ((int(code[(pos + b) >> 3]) >> ((pos + b) & 7)) & 1) << b
        return v


class CompressedIDInvertedListsWaveletTree(InvertedListsArrayCodes):
    """Wavelet-tree ids (custom_invlists_impl.h:100-124, .cpp:346-397): ONE structure over S[id] = list_no;
    get_single_id(list_no, offset) = wt.select(offset + 1, list_no). wt_type 0 = sdsl::wt_int<> (plain bit
    vectors); wt_type 1 = rrr_vector<63>: the levels as RRR(63) blocks. Sizes are those of this repository's
    wavelet matrix (bits + rank / select directories), not sdsl::size_in_bytes (SDSL absent: unpinned)."""

    def __init__(self, il: InvertedLists, wt_type: int = 0, ctx: Optional[capi.Context] = None):
        super().__init__(il)
        assert wt_type in (0, 1)  # custom_invlists_impl.cpp:349
        self.wt_type = int(wt_type)
        self.ctx = ctx or default_context()
        offsets, ids = il.csr()
        # ids must be ascending per list and < ntotal (asserts :358-359): the codec verifies it on the device
        self.blob = self.ctx.wt_encode(offsets, ids, wt_type=wt_type)
        self.codes_all = [c.copy() for c in il.codes]  # codes stay in list order (:363-364)
        self.compressed_ids_size_in_bytes = self.blob.bits_bytes + self.blob.aux_bytes  # :368,371 size_in_bytes(wt)
        self.codes_size_in_bytes = int(sum(c.size for c in self.codes_all))
        self.overhead_in_bytes = 0

    def _decode_lists(self, list_nos):
        return self.blob.decode(list_nos)

    def get_single_id(self, list_no: int, offset: int) -> int:  # :377-379
        return int(self.blob.select([list_no], [offset])[0])

    def get_single_ids(self, list_nos, offsets) -> np.ndarray:
        return self.blob.select(list_nos, offsets)


def lo_listno(label):
    return np.asarray(label, dtype=np.int64) >> 32


def lo_offset(label):
    return np.asarray(label, dtype=np.int64) & 0xFFFFFFFF


def translate_labels(invlists: InvertedListsArrayCodes, labels: np.ndarray, decode_1by1: bool = False) -> np.ndarray:
    """The id-translation half of search_IVF_defer_id_decoding (custom_invlists_impl.cpp:464-525): labels hold
    (list_no << 32 | offset) pairs from search_preassigned(store_pairs=true); negative labels pass through.
    Hits are grouped by list and only the hit lists are decoded -- in ONE bulk GPU call."""
    labels = np.asarray(labels, dtype=np.int64)
    out = labels.copy().ravel()
    valid = np.nonzero(out >= 0)[0]
    if valid.size == 0:
        return out.reshape(labels.shape)
    lists, offs = lo_listno(out[valid]), lo_offset(out[valid])
    if decode_1by1:
        if hasattr(invlists, "get_single_ids"):
            out[valid] = invlists.get_single_ids(lists, offs)
        else:
            out[valid] = [invlists.get_single_id(int(l), int(o)) for l, o in zip(lists, offs)]
        return out.reshape(labels.shape)
    if isinstance(invlists, CompressedIDInvertedListsFenwickTree):
        # one C-ABI call: distinct hit lists decoded once on the GPU, ids gathered on the device
        return invlists.blob.translate(labels.ravel()).reshape(labels.shape)
    invlists.prefetch(np.unique(lists))
    for l in np.unique(lists):
        m = lists == l
        out[valid[m]] = invlists.get_ids(int(l))[offs[m]]
    return out.reshape(labels.shape)
