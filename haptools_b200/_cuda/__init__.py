"""
ctypes binding of libhaptools_cuda.so (include/haptools_cuda.h) -- the C-ABI boundary of the
haplotype-transform hot path.  There is NO CPU fallback: if the library is missing or was
built against another ABI version, importing a symbol from here raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# HT_LIB_PATH lets experiments (profiles/) load an alternative build of the same ABI
LIB_PATH = Path(os.environ.get("HT_LIB_PATH") or Path(__file__).resolve().parent / "libhaptools_cuda.so")
ABI_VERSION = 9

HT_OK, HT_ERR_BAD_ARG, HT_ERR_CUDA, HT_ERR_UNSUPPORTED = 0, 1, 2, 3
HT_FLAG_MISSING, HT_FLAG_NON_BIALLELIC = 1, 2
HT_PACK_AUTO, HT_PACK_GENERIC, HT_PACK_VARIANT_MAJOR, HT_PACK_SAMPLE_MAJOR = 0, 1, 2, 3
HT_PACK_SAMPLE_MAJOR_BIALLELIC = 4
HT_COPY_H2D, HT_COPY_D2H, HT_COPY_D2D = 1, 2, 3
TILE_SAMPLES = 1024
RECORD_BYTES = 256

_p, _i64, _i32, _u32, _u64, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t



class PackTables(C.Structure):
    """struct ht_pack_tables (include/haptools_cuda.h); pointers are device addresses."""

    _fields_ = [
        ("n_cols", _i64), ("n_vars", _i64), ("n_keys", _i64),
        ("var_col", _p), ("var_key_off", _p), ("key_allele", _p), ("key_label", _p),
        ("col_var", _p), ("col_key01", _p), ("col_key_range", _p), ("var_other", _p), ("n_var_other", _i64),
    ]


class LocalTables(C.Structure):
    """struct ht_local_tables (include/haptools_cuda.h); pointers are device addresses."""

    _fields_ = [
        ("n_blocks", _i64), ("dmax", _i64),
        ("blk_hap_off", _p), ("blk_key_lo", _p), ("blk_key_cnt", _p),
        ("s_hap_off", _p), ("s_hap_key", _p), ("hap_order", _p),
    ]


# name -> (restype, argtypes); every symbol include/haptools_cuda.h declares
SIGNATURES = {
    "ht_version": (_i32, []),
    "ht_last_error": (C.c_char_p, []),
    "ht_num_tiles": (_i64, [_i64]),
    "ht_plane_bytes": (_sz, [_i64, _i64]),
    "ht_pack_u8": (_i32, [_p, _i64, _i64, _i64, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _i32, _p]),
    "ht_pack_i32": (_i32, [_p, _i64, _i64, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "ht_pack_i32_anc": (_i32, [_p, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p]),
    "ht_transform_u8": (_i32, [_p, _i64, _i64, _p, _p, _p, _p, _i64, _p, _i64, _p, _p]),
    "ht_transform_u8_staged": (_i32, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p]),
    "ht_transform_bits": (_i32, [_p, _i64, _i64, _p, _p, _i64, _p, _p]),
    "ht_transform_bits_gather": (_i32, [_p, _i64, _i64, _p, _p, _p, _i64, _p, _i32, _i64, _p]),
    "ht_peer_alloc": (_i32, [_sz, _p, _p]),
    "ht_peer_open": (_i32, [_p, _p]),
    "ht_peer_close": (_i32, [_p]),
    "ht_peer_free": (_i32, [_p]),
    "ht_transform_bits_local": (_i32, [_p, _i64, _i64, _p, _i64, _p, _p, _p]),
    "ht_unpack_bits_u8": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "ht_local_workspace_bytes": (_sz, [_i64, _i64]),
    "ht_transform_u8_local": (_i32, [_p, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _sz, _p]),
    "ht_unpack_bits_hapmajor": (_i32, [_p, _i64, _i64, _i64, _i64, _i32, _p, _i64, _p]),
    "ht_format_vcf_gt": (_i32, [_p, _i64, _i64, _i64, _i64, _p, _i64, _p]),
    "ht_count_bits": (_i32, [_p, _i64, _i64, _p, _p]),
    "ht_synth_gt": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _i64, _u64, _i32, _u32, _p]),
    "ht_synth_ancestry": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _i64, _u64, _u32, _u32, _p]),
    "ht_relayout_u8": (_i32, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _i64, _p]),
    "ht_bp_ancestry": (_i32, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p]),
    "ht_copy2d": (_i32, [_p, _i64, _p, _i64, _i64, _i64, _i32, _p]),
    "ht_copy_rows_h2d": (_i32, [_p, _i64, _p, _i64, _i64, _p, _i64, _p]),
    "ht_host_register": (_i32, [_p, _sz]),
    "ht_host_unregister": (_i32, [_p]),
    "ht_rowbits_pitch": (_i64, [_i64]),
    "ht_bits_to_rowbits": (_i32, [_p, _i64, _i64, _p, _i64, _p]),
    "ht_host_submit_upload": (_i64, [_p, _i64, _p, _i64, _i64, _i64, _p, _p]),
    "ht_host_submit_download": (_i64, [_p, _i64, _p, _i64, _i64, _i64, _p]),
    "ht_host_submit_download_expand": (_i64, [_p, _i64, _i64, _p, _i64, _i64, _p]),
    "ht_host_wait": (_i32, [_i64]),
    "ht_host_workers": (_i32, []),
    "ht_host_is_pageable": (_i32, [_p]),
    "ht_expand_rowbits_host": (_i32, [_p, _i64, _i64, _i64, _p, _i64]),
    "ht_host_expand_isa": (_i32, [_i32]),
    "ht_qc_u8": (_i32, [_p, _i64, _i64, _i64, _i32, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
}

_lib = None


class HaptoolsCudaError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded library (loaded once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise HaptoolsCudaError(
                f"{LIB_PATH} is missing: build it with `python -m haptools_b200._cuda.build` "
                "(there is no CPU fallback for the transform)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        got = handle.ht_version()
        if got != ABI_VERSION:
            raise HaptoolsCudaError(f"libhaptools_cuda ABI {got} != expected {ABI_VERSION}: rebuild")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    """Turn a non-zero return code into the exception class the Python API documents."""
    if rc == HT_OK:
        return
    msg = lib().ht_last_error().decode("utf-8", "replace")
    if rc == HT_ERR_BAD_ARG:
        raise ValueError(f"{what}: {msg}")
    raise HaptoolsCudaError(f"{what}: {msg} (code {rc})")
