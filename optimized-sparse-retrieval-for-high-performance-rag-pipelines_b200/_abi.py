"""ctypes binding of libb200ret.so (declared in include/b200ret.h).

The library is the product: if it is missing or does not export a declared symbol this module
raises at import time -- there is no CPU fallback and no other backend.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2R_LIB_PATH: load another build of the same library (same-box A/B measurements of two kernel variants)
LIB_PATH = os.environ.get("B2R_LIB_PATH") or os.path.join(_HERE, "libb200ret.so")

OK = 0
KIND_BM25 = 0
KIND_IMPACT = 1
TOPK_MAX_FAST = 1024


class B2RIndex(C.Structure):
    """struct b2r_index (include/b200ret.h)."""
    _fields_ = [
        ("n_docs", C.c_int64),
        ("doc_id_base", C.c_int64),
        ("nnz", C.c_int64),
        ("n_vocab", C.c_int32),
        ("tile_docs", C.c_int32),
        ("n_tiles", C.c_int32),
        ("kind", C.c_int32),
        ("post_doc", C.c_void_p),
        ("post_val", C.c_void_p),
        ("blk_ptr", C.c_void_p),
        ("dense_id", C.c_void_p),
        ("dense_ptr", C.c_void_p),
        ("n_dense_max", C.c_int32),
        ("reserved0", C.c_int32),
        ("post_pk", C.c_void_p),
    ]


class B2RIndexSizes(C.Structure):
    _fields_ = [
        ("post_doc_bytes", C.c_size_t),
        ("post_val_bytes", C.c_size_t),
        ("blk_ptr_bytes", C.c_size_t),
        ("scratch_bytes", C.c_size_t),
        ("dense_id_bytes", C.c_size_t),
        ("dense_ptr_bytes", C.c_size_t),
        ("n_dense_max", C.c_int64),
    ]


class B2RFileSection(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("bytes", C.c_uint64), ("checksum", C.c_uint64)]


SEC_NAMES = ("post_doc", "post_val", "blk_ptr", "dense_id", "dense_ptr", "idf")   # B2R_SEC_* order
FILE_ALIGN = 4096


class B2RIndexFileHeader(C.Structure):
    """struct b2r_index_file_header (include/b200ret.h): first 4096-byte page of an index file."""
    _fields_ = [
        ("magic", C.c_char * 8),
        ("version", C.c_uint32),
        ("header_bytes", C.c_uint32),
        ("n_docs", C.c_int64),
        ("doc_id_base", C.c_int64),
        ("nnz", C.c_int64),
        ("n_vocab", C.c_int32),
        ("tile_docs", C.c_int32),
        ("n_tiles", C.c_int32),
        ("kind", C.c_int32),
        ("n_dense_max", C.c_int32),
        ("subtiles", C.c_int32),
        ("k1", C.c_double),
        ("b", C.c_double),
        ("avgdl", C.c_double),
        ("sections", B2RFileSection * 6),
    ]


_P = C.c_void_p
_I32, _I64, _SZ, _F64 = C.c_int32, C.c_int64, C.c_size_t, C.c_double
_PIX = C.POINTER(B2RIndex)

# name -> (restype, argtypes); must list every symbol include/b200ret.h declares
SIGNATURES = {
    "b2r_version": (C.c_int, []),
    "b2r_last_error": (C.c_char_p, []),
    "b2r_launch_count": (C.c_ulonglong, []),
    "b2r_index_sizes_for": (C.c_int, [_I64, _I64, _I32, _I32, _I32, C.POINTER(B2RIndexSizes)]),
    "b2r_index_build": (C.c_int, [_PIX, _P, _P, _P, _P, _F64, _F64, _F64, _P, _SZ, _P]),
    "b2r_index_build_status": (C.c_int, [_P, _P]),
    "b2r_index_pack_bytes": (_SZ, [_I64]),
    "b2r_index_pack": (C.c_int, [_PIX, _P]),
    "b2r_index_pack_status": (C.c_int, [_PIX, _P, C.POINTER(C.c_float)]),
    "b2r_set_approx_prefilter": (None, [C.c_int]),
    "b2r_checksum64": (C.c_uint64, [_P, _SZ]),
    "b2r_index_file_layout": (C.c_int, [C.POINTER(B2RIndexFileHeader), C.POINTER(C.c_uint64)]),
    "b2r_index_file_check": (C.c_int, [C.POINTER(B2RIndexFileHeader), C.c_uint64]),
    "b2r_search_workspace": (C.c_int, [_PIX, _I32, _I32, C.POINTER(_SZ), C.POINTER(_SZ)]),
    "b2r_search_batch": (C.c_int, [_PIX, _P, _P, _P, _P, _I32, _I32, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "b2r_set_fused_selection": (None, [C.c_int]),
    "b2r_set_fused_cap": (None, [C.c_int]),
    "b2r_exchange_bytes": (_SZ, [_I32, _I32, _I32]),
    "b2r_exchange_merge": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "b2r_exchange_status": (C.c_int, [_P, _P]),
    "b2r_set_bank_schedule": (None, [C.c_int]),
    "b2r_set_profiling": (C.c_int, [C.c_int]),
    "b2r_profile_fused_ms": (C.c_int, [C.POINTER(C.c_float), _P]),
    "b2r_fused_plan": (C.c_int, [_PIX, _I32, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32)]),
    "b2r_search_host_extra_bytes": (_SZ, [_I32, _I64, _I32]),
    "b2r_search_batch_host": (C.c_int, [_PIX, _P, _P, _P, _P, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "b2r_topk_workspace": (C.c_int, [_I64, _I64, _I32, C.POINTER(_SZ)]),
    "b2r_topk": (C.c_int, [_P, _I64, _I64, _I64, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "b2r_merge_candidates": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "b2r_decode_keys": (C.c_int, [_P, _I64, _P, _P, _P]),
    "b2r_int8_dot_batch": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _P, _P, _P]),
    "b2r_set_int8_mma": (None, [C.c_int]),
    "b2r_set_int8_cluster": (None, [C.c_int]),
    "b2r_set_int8_pair": (None, [C.c_int]),
    "b2r_set_int8_fused": (None, [C.c_int]),
    "b2r_int8_scan_workspace": (C.c_int, [_I32, _I64, _I32, _I32, C.POINTER(_SZ)]),
    "b2r_f32_dot_topk_workspace": (C.c_int, [_I32, _I64, _I32, C.POINTER(_SZ)]),
    "b2r_f32_dot_topk": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _I64, _P, _I64, _P, _P, _P, _SZ, _P]),
    "b2r_int8_rerank_workspace": (C.c_int, [_I32, _I32, _I32, C.POINTER(_SZ)]),
    "b2r_int8_rerank": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _P, _P, _I64, _I32, _I64, _F64, _F64, _I32, _P, _P,
                                  _P, _P, _SZ, _P]),
    "b2r_int8_scan_topk": (C.c_int, [_P, _I32, _P, _I64, _I32, _P, _P, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
}


class B2RError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C <package>/csrc`). There is no CPU fallback for the retrieval hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise ImportError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    """Map a b2r_status to the reference-facing exception types."""
    if rc == OK:
        return
    msg = (lib.b2r_last_error() or b"").decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if rc in (-1, -5):          # B2R_ERR_ARG / B2R_ERR_DATA
        raise ValueError(text)
    raise B2RError(f"{text} (status {rc})")
