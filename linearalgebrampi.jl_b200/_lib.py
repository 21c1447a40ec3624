"""ctypes binding of libhpcla_b200.so (include/hpcla_b200.h).

The product path has no fallback: if the shared library is missing this module raises, it never substitutes a
CPU implementation (the CPU oracle under oracle/ is test infrastructure and is never imported from here).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPCLA_LIB: another build of the same library (kernel A/B experiments of tools/tune_spmv.py); never a fallback
LIB_PATH = os.environ.get("HPCLA_LIB") or os.path.join(_HERE, "libhpcla_b200.so")

F32, F64, C128 = 0, 1, 2
I32, I64 = 0, 1

_DTYPE_CODES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.complex128): C128}
_ITYPE_CODES = {np.dtype(np.int32): I32, np.dtype(np.int64): I64}


class HPCLAError(RuntimeError):
    pass


def dtype_code(dt) -> int:
    try:
        return _DTYPE_CODES[np.dtype(dt)]
    except KeyError:
        raise HPCLAError(f"element type {dt} is not supported (Float32, Float64, ComplexF64)") from None


def itype_code(it) -> int:
    try:
        return _ITYPE_CODES[np.dtype(it)]
    except KeyError:
        raise HPCLAError(f"index type {it} is not supported (Int32, Int64)") from None


_lib = None

# name -> (restype, argtypes); every exported symbol of include/*.h is listed here (tests check this against nm)
_i, _i64, _vp, _u64 = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_uint64
SIGNATURES = {
    "hpcla_abi_version": (_i, []),
    "hpcla_last_error": (ctypes.c_char_p, []),
    "hpcla_ctx_create": (_i, [_i, _i, _i, _vp]),
    "hpcla_nccl_unique_id": (_i, [_vp]),
    "hpcla_ctx_init_nccl": (_i, [_vp, _vp]),
    "hpcla_ctx_adopt_nccl": (_i, [_vp, _vp]),
    "hpcla_ctx_form_group": (_i, [_vp, _i]),
    "hpcla_ctx_sync": (_i, [_vp]),
    "hpcla_ctx_destroy": (None, [_vp]),
    "hpcla_host_alloc": (_i, [_vp, _i64, _vp, _vp]),
    "hpcla_host_free": (_i, [_vp, _vp]),
    "hpcla_uniform_partition": (_i, [_i64, _i, _vp]),
    "hpcla_compress_columns": (_i, [_i, _i64, _vp, _i64, _vp, _vp, _vp]),
    "hpcla_plan_begin": (_i, [_i, _i, _vp, _i64, _vp, _vp]),
    "hpcla_planb_counts": (_i, [_vp, _vp]),
    "hpcla_planb_requests": (_i, [_vp, _i, _vp]),
    "hpcla_plan_finish": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_plan_import": (_i, [_i, _i, _i, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hpcla_plan_len": (_i, [_vp, _i, _i64, _vp]),
    "hpcla_plan_get": (_i, [_vp, _i, _i64, _vp]),
    "hpcla_plan_n_gathered": (_i, [_vp, _vp]),
    "hpcla_plan_destroy": (None, [_vp]),
    "hpcla_transpose_begin": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_tb_counts": (_i, [_vp, _vp]),
    "hpcla_tb_message": (_i, [_vp, _i, _vp, _vp]),
    "hpcla_transpose_finish": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_tb_result": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "hpcla_tb_destroy": (None, [_vp]),
    "hpcla_transpose_device": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_dtb_sizes": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_dtb_result": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_dtb_destroy": (None, [_vp]),
    "hpcla_csr_create": (_i, [_vp, _i, _i, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "hpcla_csr_info": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_csr_tile_classes": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "hpcla_csr_destroy": (None, [_vp]),
    "hpcla_spmv_create": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "hpcla_spmv_run": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_spmv_run_staged": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_spmv_graph_capture": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_spmv_graph_launch": (_i, [_vp, _vp]),
    "hpcla_spmv_timeline": (_i, [_vp, _vp]),
    "hpcla_spmv_halo_blob_size": (_i, [_vp, _vp]),
    "hpcla_spmv_halo_export": (_i, [_vp, _vp]),
    "hpcla_spmv_halo_connect": (_i, [_vp, _vp]),
    "hpcla_spmv_halo_debug": (_i, [_vp, _vp]),
    "hpcla_spmv_begin": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_spmv_finish": (_i, [_vp]),
    "hpcla_spmm_run": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _vp]),
    "hpcla_spmm_begin": (_i, [_vp, _vp, _i64, _vp, _i64, _i, _vp]),
    "hpcla_spmm_finish": (_i, [_vp]),
    "hpcla_spmv_gather": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_spmv_gather_finish": (_i, [_vp]),
    "hpcla_spmv_info": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "hpcla_spmv_tile_lists": (_i, [_vp, _vp]),
    "hpcla_spmv_launch_count": (_i64, [_vp]),
    "hpcla_spmv_destroy": (None, [_vp]),
    "hpcla_spgemm_symbolic": (_i, [_i, _i64, _vp, _vp, _i64, _vp, _vp, _vp]),
    "hpcla_spgemm_sizes": (_i, [_vp, _vp, _vp, _vp]),
    "hpcla_spgemm_structure": (_i, [_vp, _i, _vp, _vp, _vp]),
    "hpcla_spgemm_numeric": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "hpcla_spgemm_destroy": (None, [_vp]),
    "hpcla_exchange_bytes": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_dot": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "hpcla_nrm2": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "hpcla_axpby": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_cg": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "hpcla_repartition_plan": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hpcla_repartition_run": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}


def lib() -> ctypes.CDLL:
    """The loaded shared library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HPCLAError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for this backend."
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.hpcla_abi_version() != 1:
            raise HPCLAError("libhpcla_b200.so has an unexpected ABI version; rebuild it")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().hpcla_last_error()
        raise HPCLAError(f"[hpcla status {rc}] {msg.decode() if msg else 'unknown error'}")


def ptr(a) -> int:
    """Address of a numpy array / torch tensor (0 for None or empty)."""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data if a.size else 0
    return a.data_ptr() if a.numel() else 0  # torch tensor


def ptr_array(arrs):
    return (ctypes.c_void_p * max(len(arrs), 1))(*[ptr(a) or None for a in arrs])
