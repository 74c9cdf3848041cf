"""Thin Python front-end of the C ABI (include/idcodec.h).

Inputs may be numpy arrays (host memory) or torch CUDA tensors (device memory);
results come back in the same kind of memory. Nothing here computes: every call
goes straight into libidcodec.so, which launches the sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import EfInfo, RocInfo, WtInfo

MEM_HOST, MEM_DEVICE = 0, 1
F_SORTED, F_PRECISION_SAFE, F_WANT_ORDER = 1, 2, 4


class IdcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"idcodec error {code}: {msg}")
        self.code = code


def _check(rc: int) -> None:
    if rc != 0:
        raise IdcError(rc, _lib.load().idc_last_error().decode())


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> tuple[Optional[int], int]:
    """-> (address, mem kind) of a numpy array or torch tensor (None -> NULL)."""
    if x is None:
        return None, MEM_HOST
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr(), (MEM_DEVICE if x.is_cuda else MEM_HOST)
    assert isinstance(x, np.ndarray) and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data, MEM_HOST


def _host_u64(a) -> np.ndarray:
    if _is_torch(a):
        a = a.cpu().numpy()
    return np.ascontiguousarray(np.asarray(a).astype(np.uint64, copy=False))


def _ids_array(ids):
    """ids as int64/int32 contiguous; keeps torch CUDA tensors on the device."""
    if _is_torch(ids):
        import torch

        assert ids.dtype in (torch.int64, torch.int32)
        return ids.contiguous(), (8 if ids.dtype == torch.int64 else 4)
    ids = np.asarray(ids)
    if ids.dtype in (np.int32, np.uint32):
        return np.ascontiguousarray(ids), 4
    return np.ascontiguousarray(ids.astype(np.int64, copy=False)), 8


class Context:
    """One codec context = one device + one stream (idc_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._l = _lib.load()
        h = C.c_void_p()
        if stream is None:
            _check(self._l.idc_ctx_create(device, C.byref(h)))
        else:
            _check(self._l.idc_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        self._h = h
        self.device = device

    def close(self) -> None:
        """Destroy the context. Raises (and keeps the context) while blobs created by it are alive: they hand their
        device arrays back to the context's pool when they are freed (idc_ctx_destroy refuses, include/idcodec.h)."""
        if getattr(self, "_h", None):
            _check(self._l.idc_ctx_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self) -> None:
        _check(self._l.idc_ctx_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._l.idc_ctx_launch_count(self._h))

    def set_timing(self, on: bool) -> None:
        _check(self._l.idc_ctx_set_timing(self._h, 1 if on else 0))

    def last_kernel_ms(self) -> float:
        return float(self._l.idc_ctx_last_kernel_ms(self._h))

    def last_kernel_breakdown(self) -> list[tuple[str, float]]:
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = self._l.idc_ctx_last_kernel_breakdown(self._h, names, ms, 64)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    # ------------------------------------------------------------- ROC ----
    def roc_encode(self, offsets, ids, *, sorted_ids: bool = False, precision_safe: bool = False,
                   want_order: bool = False, max_unit: int = 65536) -> "RocBlob":
        offsets = _host_u64(offsets)
        ids, idb = _ids_array(ids)
        p, mem = _ptr(ids)
        flags = (F_SORTED if sorted_ids else 0) | (F_PRECISION_SAFE if precision_safe else 0) | (
            F_WANT_ORDER if want_order else 0)
        h = C.c_void_p()
        _check(self._l.idc_roc_encode(self._h, offsets.size - 1, offsets.ctypes.data, p, idb, mem, flags, max_unit,
                                      C.byref(h)))
        return RocBlob(self, h)

    def roc_encode_rows(self, data, *, precision_safe: bool = False, want_order: bool = False) -> "RocBlob":
        if not _is_torch(data):
            data = np.ascontiguousarray(data, dtype=np.int32)
        n, k = data.shape
        p, mem = _ptr(data)
        flags = (F_PRECISION_SAFE if precision_safe else 0) | (F_WANT_ORDER if want_order else 0)
        h = C.c_void_p()
        _check(self._l.idc_roc_encode_rows(self._h, n, k, p, mem, flags, C.byref(h)))
        return RocBlob(self, h)

    def roc_import(self, unit_n, precision, heads, word_offsets, words) -> "RocBlob":
        unit_n = np.ascontiguousarray(unit_n, dtype=np.uint32)
        precision = np.ascontiguousarray(precision, dtype=np.uint8)
        heads = np.ascontiguousarray(heads, dtype=np.uint64)
        word_offsets = np.ascontiguousarray(word_offsets, dtype=np.uint64)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        if words.size == 0:
            words = np.zeros(1, np.uint32)
        h = C.c_void_p()
        _check(self._l.idc_roc_blob_import(self._h, unit_n.size, unit_n.ctypes.data, precision.ctypes.data,
                                           heads.ctypes.data, word_offsets.ctypes.data, words.ctypes.data, C.byref(h)))
        return RocBlob(self, h)

    def roc_assemble(self, list_offsets, payload: dict, *, max_unit: int = 65536) -> "RocBlob":
        """idc_roc_blob_assemble: a blob over the lists of `list_offsets` from per-unit payload arrays in unit order
        (numpy, or torch tensors on the host or on this context's device; all six from the same place)."""
        off = _host_u64(list_offsets)
        arrs, mems = [], set()
        for k in RocBlob.PAYLOAD_KEYS:
            a = payload.get(k)
            if a is None:
                arrs.append(None)
                continue
            if not _is_torch(a):
                a = np.ascontiguousarray(a)
            ptr, mem = _ptr(a)
            arrs.append((a, ptr))
            mems.add(mem)
        assert len(mems) <= 1, "payload arrays must all live in the same memory"
        mem = mems.pop() if mems else MEM_HOST
        nwords_total = int(payload["words"].shape[0]) if payload.get("words") is not None else 0
        h = C.c_void_p()
        _check(self._l.idc_roc_blob_assemble(self._h, off.size - 1, off.ctypes.data, max_unit, mem,
                                             *((a[1] if a is not None else None) for a in arrs), nwords_total, C.byref(h)))
        return RocBlob(self, h)

    def roc_load(self, path) -> "RocBlob":
        """idc_roc_blob_load: a blob from its flat file form (RocBlob.save)."""
        h = C.c_void_p()
        _check(self._l.idc_roc_blob_load(self._h, str(path).encode(), C.byref(h)))
        return RocBlob(self, h)

    # ------------------------------------------------------------- EF -----
    def ef_import(self, list_offsets, universe, low, high, *, row_stride: int = 0) -> "EfBlob":
        """idc_ef_blob_import: the inverse of EfBlob.export (bit vectors as numpy arrays or torch tensors, both on the
        host or both on this context's device); samples and chunk directory are rebuilt on the device."""
        off = _host_u64(list_offsets)
        uni = np.ascontiguousarray(universe, dtype=np.uint64)
        if not _is_torch(low):
            low = np.ascontiguousarray(low, dtype=np.uint64)
            high = np.ascontiguousarray(high, dtype=np.uint64)
        pl, mem = _ptr(low)
        ph, mem2 = _ptr(high)
        assert mem == mem2, "low and high must live in the same memory"
        h = C.c_void_p()
        _check(self._l.idc_ef_blob_import(self._h, off.size - 1, off.ctypes.data, uni.ctypes.data, int(row_stride), pl, ph, mem, C.byref(h)))
        return EfBlob(self, h)

    def ef_load(self, path) -> "EfBlob":
        h = C.c_void_p()
        _check(self._l.idc_ef_blob_load(self._h, str(path).encode(), C.byref(h)))
        return EfBlob(self, h)

    def ef_encode(self, offsets, ids, *, sorted_ids: bool = False) -> "EfBlob":
        offsets = _host_u64(offsets)
        ids, idb = _ids_array(ids)
        p, mem = _ptr(ids)
        h = C.c_void_p()
        _check(self._l.idc_ef_encode(self._h, offsets.size - 1, offsets.ctypes.data, p, idb, mem,
                                     F_SORTED if sorted_ids else 0, C.byref(h)))
        return EfBlob(self, h)

    def ef_encode_rows(self, data) -> "EfBlob":
        if not _is_torch(data):
            data = np.ascontiguousarray(data, dtype=np.int32)
        n, k = data.shape
        p, mem = _ptr(data)
        h = C.c_void_p()
        _check(self._l.idc_ef_encode_rows(self._h, n, k, p, mem, 0, C.byref(h)))
        return EfBlob(self, h)

    # ---------------------------------------------------- wavelet tree ----
    def wt_import(self, exported: dict, *, wt_type: int = 0) -> "WtBlob":
        """idc_wt_blob_import: the inverse of WtBlob.export (host arrays)."""
        off = _host_u64(exported["list_offsets"])
        arrs = [np.ascontiguousarray(exported[k], dtype=t) for k, t in (("bits", np.uint64), ("rank", np.uint32), ("sel1", np.uint32),
                                                                         ("sel0", np.uint32), ("start", np.uint32))]
        h = C.c_void_p()
        _check(self._l.idc_wt_blob_import(self._h, off.size - 1, off.ctypes.data, int(wt_type), *(a.ctypes.data for a in arrs), MEM_HOST,
                                          C.byref(h)))
        return WtBlob(self, h)

    def wt_load(self, path) -> "WtBlob":
        h = C.c_void_p()
        _check(self._l.idc_wt_blob_load(self._h, str(path).encode(), C.byref(h)))
        return WtBlob(self, h)

    def wt_encode(self, offsets, ids, *, wt_type: int = 0) -> "WtBlob":
        offsets = _host_u64(offsets)
        ids, idb = _ids_array(ids)
        p, mem = _ptr(ids)
        h = C.c_void_p()
        _check(self._l.idc_wt_encode(self._h, offsets.size - 1, offsets.ctypes.data, p, idb, mem, int(wt_type),
                                     C.byref(h)))
        return WtBlob(self, h)

    # ------------------------------------------------------------ bits ----
    def bits_pack(self, vals, bits: int, nbytes: Optional[int] = None):
        n = int(vals.shape[0])
        if nbytes is None:
            nbytes = (n * bits + 7) // 8
        if _is_torch(vals):
            import torch

            vb = 8 if vals.dtype == torch.int64 else 4
            out = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=vals.device)
        else:
            vals = np.ascontiguousarray(vals)
            vb = vals.dtype.itemsize
            out = np.empty(max(nbytes, 1), dtype=np.uint8)
        pv, mv = _ptr(vals)
        po, mo = _ptr(out)
        _check(self._l.idc_bits_pack(self._h, n, pv, vb, mv, bits, po, nbytes, mo))
        return out[:nbytes]

    def bits_unpack(self, code, n: int, bits: int, val_bytes: int = 8):
        if _is_torch(code):
            import torch

            out = torch.empty(max(n, 1), dtype=torch.int64 if val_bytes == 8 else torch.int32, device=code.device)
        else:
            code = np.ascontiguousarray(code, dtype=np.uint8)
            out = np.empty(max(n, 1), dtype=np.uint64 if val_bytes == 8 else np.uint32)
        pc, mc = _ptr(code)
        po, mo = _ptr(out)
        _check(self._l.idc_bits_unpack(self._h, n, pc, int(code.shape[0]), mc, bits, po, val_bytes, mo))
        return out[:n]


def _alloc_like(device_tensor_or_none, n: int, dtype_np, torch_device=None):
    if torch_device is not None:
        import torch

        tdt = {np.int64: torch.int64, np.int32: torch.int32, np.uint32: torch.int32}[dtype_np]
        return torch.empty(max(n, 1), dtype=tdt, device=torch_device)
    return np.empty(max(n, 1), dtype=dtype_np)


class RocBlob:
    """Device-resident ROC blob (idc_roc_blob)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._l = ctx._l
        self._h = handle
        info = RocInfo()
        _check(self._l.idc_roc_blob_info(self._h, C.byref(info)))
        self.info = info

    def free(self) -> None:
        if getattr(self, "_h", None):
            self._l.idc_roc_blob_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    nlist = property(lambda s: int(s.info.nlist))
    nunits = property(lambda s: int(s.info.nunits))
    total_ids = property(lambda s: int(s.info.total_ids))
    total_words = property(lambda s: int(s.info.total_words))
    ans_bytes = property(lambda s: int(s.info.ans_bytes))
    row_stride = property(lambda s: int(s.info.row_stride))

    def save(self, path) -> None:
        """idc_roc_blob_save: the blob's flat file form (csrc/idc_file.h); Context.roc_load reads it back."""
        _check(self._l.idc_roc_blob_save(self._h, str(path).encode()))

    def export(self) -> dict:
        i = self.info
        d = dict(
            list_offsets=np.zeros(i.nlist + 1, np.uint64), unit_offsets=np.zeros(i.nlist + 1, np.uint64),
            unit_n=np.zeros(max(i.nunits, 1), np.uint32), precision=np.zeros(max(i.nunits, 1), np.uint8),
            heads=np.zeros(max(i.nunits, 1), np.uint64), word_offsets=np.zeros(i.nunits + 1, np.uint64),
            words=np.zeros(max(i.total_words, 1), np.uint32),
        )
        _check(self._l.idc_roc_blob_export(self._h, *(d[k].ctypes.data for k in (
            "list_offsets", "unit_offsets", "unit_n", "precision", "heads", "word_offsets", "words"))))
        for k in ("unit_n", "precision", "heads"):
            d[k] = d[k][: i.nunits]
        d["words"] = d["words"][: i.total_words]
        return d

    PAYLOAD_KEYS = ("precision", "heads", "nwords", "lo", "hi", "words")

    def export_payload(self, device=None) -> dict:
        """The wire form of the blob (idc_roc_blob_export_payload): per-unit precision / head / word count / id-range
        hints plus all stream words, as torch tensors on `device` (device -> device copies, nothing touches the
        host) or, with device=None, as numpy arrays. Unsigned arrays travel as same-width signed torch dtypes."""
        i = self.info
        nu, nw = int(i.nunits), int(i.total_words)
        if device is not None:
            import torch

            mk = lambda n, dt: torch.empty(max(n, 1), dtype=dt, device=device)
            d = dict(precision=mk(nu, torch.uint8), heads=mk(nu, torch.int64), nwords=mk(nu, torch.int32),
                     lo=mk(nu, torch.int32), hi=mk(nu, torch.int32), words=mk(nw, torch.int32))
            ptr = lambda t: t.data_ptr()
            mem = MEM_DEVICE if torch.device(device).type == "cuda" else MEM_HOST
        else:
            mk = lambda n, dt: np.empty(max(n, 1), dtype=dt)
            d = dict(precision=mk(nu, np.uint8), heads=mk(nu, np.uint64), nwords=mk(nu, np.uint32),
                     lo=mk(nu, np.uint32), hi=mk(nu, np.uint32), words=mk(nw, np.uint32))
            ptr = lambda a: a.ctypes.data
            mem = MEM_HOST
        _check(self._l.idc_roc_blob_export_payload(self._h, mem, *(ptr(d[k]) for k in self.PAYLOAD_KEYS)))
        for k in self.PAYLOAD_KEYS:
            d[k] = d[k][: (nw if k == "words" else nu)]
        return d

    def order(self, device=None):
        n = self.nlist * self.row_stride if self.row_stride else self.total_ids
        out = _alloc_like(None, n, np.uint32, device)
        p, mem = _ptr(out)
        _check(self._l.idc_roc_blob_order(self._h, p, mem))
        return out[:n]

    def decode(self, list_nos: Optional[Sequence[int]] = None, *, id_bytes: int = 8, device=None, out=None):
        """-> (ids, out_offsets). device=None: numpy result; else a torch device. out: the caller's own buffer (numpy
        array or torch tensor of int64 / int32, e.g. pinned host memory -- the large-output path then sends the
        longest units' ids to the host while their chains are still running)."""
        if list_nos is None:
            nsel, lp = self.nlist, None
            total = self.total_ids
        else:
            ln = _host_u64(list_nos)
            nsel, lp = ln.size, ln.ctypes.data
            total = None
        out_off = np.zeros(nsel + 1, np.uint64)
        if total is None:
            ex_off = self.export_list_offsets()
            total = int(sum(int(ex_off[int(l) + 1] - ex_off[int(l)]) for l in ln))
        if out is None:
            out = _alloc_like(None, total, np.int64 if id_bytes == 8 else np.int32, device)
        else:
            isz = out.element_size() if _is_torch(out) else out.dtype.itemsize
            cnt = int(out.numel() if _is_torch(out) else out.size)
            if isz != id_bytes or cnt < total:
                raise ValueError("out: wrong element size or too small")
        p, mem = _ptr(out)
        _check(self._l.idc_roc_decode(self.ctx._h, self._h, lp, nsel, p, id_bytes, mem, out_off.ctypes.data))
        return out[:total], out_off

    def translate(self, labels, *, device=None):
        """(list_no << 32 | offset) labels -> ids; negative labels pass through (idc_roc_translate)."""
        if not _is_torch(labels):
            labels = np.ascontiguousarray(labels, dtype=np.int64)
        n = int(labels.numel() if _is_torch(labels) else labels.size)
        out = _alloc_like(None, n, np.int64, device)
        pl, ml = _ptr(labels)
        po, mo = _ptr(out)
        _check(self._l.idc_roc_translate(self.ctx._h, self._h, pl, ml, n, po, mo))
        return out[:n]

    def export_list_offsets(self) -> np.ndarray:
        lo = np.zeros(self.nlist + 1, np.uint64)
        _check(self._l.idc_roc_blob_export(self._h, lo.ctypes.data, None, None, None, None, None, None))
        return lo

    def decode_rows(self, row_nos=None, *, device=None):
        """-> (neighbors [nsel, K] int32 padded with -1, counts [nsel])."""
        K = self.row_stride
        if row_nos is None:
            nsel, rp, rmem = self.nlist, None, MEM_HOST
        else:
            if not _is_torch(row_nos):
                row_nos = np.ascontiguousarray(row_nos, dtype=np.int32)
            nsel = int(row_nos.shape[0])
            rp, rmem = _ptr(row_nos)
        out = _alloc_like(None, nsel * K, np.int32, device)
        cnt = _alloc_like(None, nsel, np.uint32, device)
        po, mo = _ptr(out)
        pc, _ = _ptr(cnt)
        _check(self._l.idc_roc_decode_rows(self.ctx._h, self._h, rp, rmem, nsel, po, pc, mo))
        return out[: nsel * K].reshape(nsel, K), cnt[:nsel]


class EfBlob:
    """Device-resident Elias-Fano blob (idc_ef_blob)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._l = ctx._l
        self._h = handle
        info = EfInfo()
        _check(self._l.idc_ef_blob_info(self._h, C.byref(info)))
        self.info = info

    def free(self) -> None:
        if getattr(self, "_h", None):
            self._l.idc_ef_blob_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    nlist = property(lambda s: int(s.info.nlist))
    total_ids = property(lambda s: int(s.info.total_ids))
    bits_total = property(lambda s: int(s.info.bits_total))
    row_stride = property(lambda s: int(s.info.row_stride))

    def save(self, path) -> None:
        """idc_ef_blob_save: the blob's flat file form (csrc/idc_file.h); Context.ef_load reads it back."""
        _check(self._l.idc_ef_blob_save(self._h, str(path).encode()))

    def export(self) -> dict:
        i = self.info
        d = dict(
            list_offsets=np.zeros(i.nlist + 1, np.uint64), l=np.zeros(max(i.nlist, 1), np.uint8),
            universe=np.zeros(max(i.nlist, 1), np.uint64), low_offsets=np.zeros(i.nlist + 1, np.uint64),
            high_offsets=np.zeros(i.nlist + 1, np.uint64), low=np.zeros(max(i.low_words, 1), np.uint64),
            high=np.zeros(max(i.high_words, 1), np.uint64),
        )
        _check(self._l.idc_ef_blob_export(self._h, *(d[k].ctypes.data for k in (
            "list_offsets", "l", "universe", "low_offsets", "high_offsets", "low", "high"))))
        d["l"] = d["l"][: i.nlist]
        d["universe"] = d["universe"][: i.nlist]
        d["low"] = d["low"][: i.low_words]
        d["high"] = d["high"][: i.high_words]
        return d

    def decode(self, list_nos=None, *, id_bytes: int = 8, device=None):
        if list_nos is None:
            nsel, lp, total = self.nlist, None, self.total_ids
        else:
            ln = _host_u64(list_nos)
            nsel, lp = ln.size, ln.ctypes.data
            lo = np.zeros(self.nlist + 1, np.uint64)
            _check(self._l.idc_ef_blob_export(self._h, lo.ctypes.data, None, None, None, None, None, None))
            total = int(sum(int(lo[int(l) + 1] - lo[int(l)]) for l in ln))
        out_off = np.zeros(nsel + 1, np.uint64)
        out = _alloc_like(None, total, np.int64 if id_bytes == 8 else np.int32, device)
        p, mem = _ptr(out)
        _check(self._l.idc_ef_decode(self.ctx._h, self._h, lp, nsel, p, id_bytes, mem, out_off.ctypes.data))
        return out[:total], out_off

    def decode_rows(self, row_nos=None, *, device=None):
        K = self.row_stride
        if row_nos is None:
            nsel, rp, rmem = self.nlist, None, MEM_HOST
        else:
            if not _is_torch(row_nos):
                row_nos = np.ascontiguousarray(row_nos, dtype=np.int32)
            nsel = int(row_nos.shape[0])
            rp, rmem = _ptr(row_nos)
        out = _alloc_like(None, nsel * K, np.int32, device)
        cnt = _alloc_like(None, nsel, np.uint32, device)
        po, mo = _ptr(out)
        pc, _ = _ptr(cnt)
        _check(self._l.idc_ef_decode_rows(self.ctx._h, self._h, rp, rmem, nsel, po, pc, mo))
        return out[: nsel * K].reshape(nsel, K), cnt[:nsel]

    def select(self, list_nos, offsets_in_list, *, device=None):
        if device is not None:
            import torch

            ql = torch.as_tensor(list_nos, dtype=torch.int64, device=device).contiguous()
            qo = torch.as_tensor(offsets_in_list, dtype=torch.int64, device=device).contiguous()
            out = torch.empty(max(ql.numel(), 1), dtype=torch.int64, device=device)
            nq = ql.numel()
        else:
            ql = _host_u64(list_nos)
            qo = _host_u64(offsets_in_list)
            out = np.empty(max(ql.size, 1), np.int64)
            nq = ql.size
        pl, mq = _ptr(ql)
        pq, _ = _ptr(qo)
        po, mo = _ptr(out)
        _check(self._l.idc_ef_select(self.ctx._h, self._h, pl, pq, nq, mq, po, mo))
        return out[:nq]



class WtBlob:
    """Device-resident wavelet structure over S[id] = list_no (idc_wt_blob)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._l = ctx._l
        self._h = handle
        info = WtInfo()
        _check(self._l.idc_wt_blob_info(self._h, C.byref(info)))
        self.info = info

    def free(self) -> None:
        if getattr(self, "_h", None):
            self._l.idc_wt_blob_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    nlist = property(lambda s: int(s.info.nlist))
    total_ids = property(lambda s: int(s.info.total_ids))
    levels = property(lambda s: int(s.info.levels))
    bits_bytes = property(lambda s: int(s.info.bits_bytes))
    aux_bytes = property(lambda s: int(s.info.aux_bytes))

    def save(self, path) -> None:
        """idc_wt_blob_save: the blob's flat file form (csrc/idc_file.h); Context.wt_load reads it back."""
        _check(self._l.idc_wt_blob_save(self._h, str(path).encode()))

    def export(self) -> dict:
        n, levels, nlist = self.total_ids, self.levels, self.nlist
        nblk = (n + 511) // 512
        samp = (n >> 11) + 2
        d = dict(
            list_offsets=np.zeros(nlist + 1, np.uint64), bits=np.zeros((levels, max(nblk * 8, 1)), np.uint64),
            rank=np.zeros((levels, nblk + 1), np.uint32), sel1=np.zeros((levels, samp), np.uint32),
            sel0=np.zeros((levels, samp), np.uint32), start=np.zeros(max(nlist, 1), np.uint32),
        )
        if nblk == 0:
            d["bits"] = np.zeros((levels, 0), np.uint64)
            _check(self._l.idc_wt_blob_export(self._h, d["list_offsets"].ctypes.data, None, None, None, None, None))
        else:
            _check(self._l.idc_wt_blob_export(self._h, *(d[k].ctypes.data for k in (
                "list_offsets", "bits", "rank", "sel1", "sel0", "start"))))
        d["start"] = d["start"][:nlist]
        d.update(levels=levels, n=n, nblk=nblk, nlist=nlist)
        return d

    def decode(self, list_nos=None, *, id_bytes: int = 8, device=None):
        if list_nos is None:
            nsel, lp, total = self.nlist, None, self.total_ids
        else:
            ln = _host_u64(list_nos)
            nsel, lp = ln.size, ln.ctypes.data
            lo = np.zeros(self.nlist + 1, np.uint64)
            _check(self._l.idc_wt_blob_export(self._h, lo.ctypes.data, None, None, None, None, None))
            total = int(sum(int(lo[int(l) + 1] - lo[int(l)]) for l in ln))
        out_off = np.zeros(nsel + 1, np.uint64)
        out = _alloc_like(None, total, np.int64 if id_bytes == 8 else np.int32, device)
        p, mem = _ptr(out)
        _check(self._l.idc_wt_decode(self.ctx._h, self._h, lp, nsel, p, id_bytes, mem, out_off.ctypes.data))
        return out[:total], out_off

    def export_rrr(self) -> dict:
        """wt_type = 1: the compressed arrays as they lie in HBM (idc_wt_blob_export_rrr)."""
        levels, nblk = self.levels, (self.total_ids + 511) // 512
        off_base = np.zeros(levels + 1, np.uint64)
        _check(self._l.idc_wt_blob_export_rrr(self._h, None, None, off_base.ctypes.data, None))
        cls = np.zeros((levels, max(nblk, 1)), np.uint64)
        ptr = np.zeros((levels, nblk + 1), np.uint32)
        off = np.zeros(max(int(off_base[-1]), 1), np.uint64)
        _check(self._l.idc_wt_blob_export_rrr(self._h, cls.ctypes.data, ptr.ctypes.data, None, off.ctypes.data))
        return dict(cls=cls[:, :nblk], ptr=ptr, off_base=off_base, off=off[: int(off_base[-1])])

    def select(self, list_nos, offsets_in_list, *, device=None):
        if device is not None:
            import torch

            ql = torch.as_tensor(list_nos, dtype=torch.int64, device=device).contiguous()
            qo = torch.as_tensor(offsets_in_list, dtype=torch.int64, device=device).contiguous()
            out = torch.empty(max(ql.numel(), 1), dtype=torch.int64, device=device)
            nq = ql.numel()
        else:
            ql = _host_u64(list_nos)
            qo = _host_u64(offsets_in_list)
            out = np.empty(max(ql.size, 1), np.int64)
            nq = ql.size
        pl, mq = _ptr(ql)
        pq, _ = _ptr(qo)
        po, mo = _ptr(out)
        _check(self._l.idc_wt_select(self.ctx._h, self._h, pl, pq, nq, mq, po, mo))
        return out[:nq]
