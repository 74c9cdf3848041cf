"""ctypes loader for libidcodec.so. Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libidcodec.so"

_lib = None

u64 = C.c_uint64
u32 = C.c_uint32
vp = C.c_void_p


class RocInfo(C.Structure):
    _fields_ = [("nlist", u64), ("nunits", u64), ("total_ids", u64), ("total_words", u64), ("ans_bytes", u64),
                ("device_bytes", u64), ("max_unit", u32), ("row_stride", u32)]


class EfInfo(C.Structure):
    _fields_ = [("nlist", u64), ("total_ids", u64), ("low_words", u64), ("high_words", u64), ("bits_total", u64),
                ("device_bytes", u64), ("row_stride", u32)]


class WtInfo(C.Structure):
    _fields_ = [("nlist", u64), ("total_ids", u64), ("bits_bytes", u64), ("aux_bytes", u64), ("device_bytes", u64),
                ("levels", u32), ("wt_type", u32)]


# every symbol include/idcodec.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "idc_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "idc_ctx_create_on_stream": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "idc_ctx_destroy": (C.c_int, [vp]),
    "idc_ctx_synchronize": (C.c_int, [vp]),
    "idc_ctx_launch_count": (u64, [vp]),
    "idc_ctx_set_timing": (C.c_int, [vp, C.c_int]),
    "idc_ctx_last_kernel_ms": (C.c_float, [vp]),
    "idc_ctx_last_kernel_breakdown": (C.c_int, [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]),
    "idc_last_error": (C.c_char_p, []),
    "idc_version": (C.c_int, []),
    "idc_roc_encode": (C.c_int, [vp, u64, vp, vp, C.c_int, C.c_int, u32, u32, C.POINTER(vp)]),
    "idc_roc_encode_rows": (C.c_int, [vp, u64, u32, vp, C.c_int, u32, C.POINTER(vp)]),
    "idc_roc_blob_info": (C.c_int, [vp, C.POINTER(RocInfo)]),
    "idc_roc_blob_export": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "idc_roc_blob_import": (C.c_int, [vp, u64, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "idc_roc_blob_export_payload": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp]),
    "idc_roc_blob_assemble": (C.c_int, [vp, u64, vp, u32, C.c_int, vp, vp, vp, vp, vp, vp, u64, C.POINTER(vp)]),
    "idc_roc_blob_order": (C.c_int, [vp, vp, C.c_int]),
    "idc_roc_blob_free": (C.c_int, [vp]),
    "idc_roc_blob_save": (C.c_int, [vp, C.c_char_p]),
    "idc_roc_blob_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "idc_roc_decode": (C.c_int, [vp, vp, vp, u64, vp, C.c_int, C.c_int, vp]),
    "idc_roc_translate": (C.c_int, [vp, vp, vp, C.c_int, u64, vp, C.c_int]),
    "idc_roc_decode_rows": (C.c_int, [vp, vp, vp, C.c_int, u64, vp, vp, C.c_int]),
    "idc_ef_encode": (C.c_int, [vp, u64, vp, vp, C.c_int, C.c_int, u32, C.POINTER(vp)]),
    "idc_ef_encode_rows": (C.c_int, [vp, u64, u32, vp, C.c_int, u32, C.POINTER(vp)]),
    "idc_ef_blob_info": (C.c_int, [vp, C.POINTER(EfInfo)]),
    "idc_ef_blob_export": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "idc_ef_blob_import": (C.c_int, [vp, u64, vp, vp, u32, vp, vp, C.c_int, C.POINTER(vp)]),
    "idc_ef_blob_save": (C.c_int, [vp, C.c_char_p]),
    "idc_ef_blob_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "idc_ef_blob_free": (C.c_int, [vp]),
    "idc_ef_decode": (C.c_int, [vp, vp, vp, u64, vp, C.c_int, C.c_int, vp]),
    "idc_ef_decode_rows": (C.c_int, [vp, vp, vp, C.c_int, u64, vp, vp, C.c_int]),
    "idc_ef_select": (C.c_int, [vp, vp, vp, vp, u64, C.c_int, vp, C.c_int]),
    "idc_wt_encode": (C.c_int, [vp, u64, vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "idc_wt_blob_info": (C.c_int, [vp, C.POINTER(WtInfo)]),
    "idc_wt_blob_export": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "idc_wt_blob_import": (C.c_int, [vp, u64, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.POINTER(vp)]),
    "idc_wt_blob_export_rrr": (C.c_int, [vp, vp, vp, vp, vp]),
    "idc_wt_blob_save": (C.c_int, [vp, C.c_char_p]),
    "idc_wt_blob_load": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
    "idc_wt_blob_free": (C.c_int, [vp]),
    "idc_wt_select": (C.c_int, [vp, vp, vp, vp, u64, C.c_int, vp, C.c_int]),
    "idc_wt_decode": (C.c_int, [vp, vp, vp, u64, vp, C.c_int, C.c_int, vp]),
    "idc_bits_pack": (C.c_int, [vp, u64, vp, C.c_int, C.c_int, C.c_int, vp, u64, C.c_int]),
    "idc_bits_unpack": (C.c_int, [vp, u64, vp, u64, C.c_int, C.c_int, vp, C.c_int, C.c_int]),
}


def load() -> C.CDLL:
    """Load libidcodec.so and bind every declared symbol; raise if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m vector_db_id_compression_b200.build` "
            "(this package has no CPU fallback)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
