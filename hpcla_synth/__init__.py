"""hpcla_synth — deterministic synthetic inputs, ctypes front-end of include/hpcla_synth.h.

Test / benchmark infrastructure in a library of its own (hpcla_synth/libhpcla_synth.so, plain g++): the GPU path, the
CPU oracle and the CPU reference arm all generate their matrices and vectors here, so they see identical bits, and the
reference arm never has to map the product library.  Nothing in this package touches CUDA or the product.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhpcla_synth.so")
SRC = os.path.join(_HERE, "hpcla_synth.cpp")
HEADER = os.path.join(_HERE, "..", "include", "hpcla_synth.h")

LAPLACE2D_5PT, POISSON3D_7PT, STENCIL3D_27PT = 0, 1, 2
X_SEED = 0x5EED
POWERLAW_SEED = 0xC4
POWERLAW_MAX_LEN = 1_000_000

_DTYPE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex128): 2}
_ITYPE = {np.dtype(np.int32): 0, np.dtype(np.int64): 1}

_i, _i64, _vp, _u64 = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_uint64
SIGNATURES = {
    "hpcla_synth_stencil_rows": (_i64, [_i, _i64, _i64, _i64]),
    "hpcla_synth_stencil_nnz": (_i64, [_i, _i64, _i64, _i64, _i64, _i64]),
    "hpcla_synth_stencil_fill": (_i, [_i, _i64, _i64, _i64, _i, _i, _i64, _i64, _vp, _vp, _vp]),
    "hpcla_synth_powerlaw_nnz": (_i64, [_i64, _u64, _i64, _i64, _i64]),
    "hpcla_synth_powerlaw_fill": (_i, [_i64, _u64, _i64, _i, _i, _i64, _i64, _vp, _vp, _vp]),
    "hpcla_synth_vector": (_i, [_i, _u64, _i64, _i64, _vp]),
    "hpcla_synth_last_error": (ctypes.c_char_p, []),
    "hpcla_synth_set_threads": (None, [_i]),
}


def build(force: bool = False) -> str:
    deps = [SRC, HEADER]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-Wall", "-pthread", "-shared", "-o", LIB_PATH, SRC])
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"hpcla_synth: {lib().hpcla_synth_last_error().decode()}")


def _p(a: np.ndarray) -> int:
    return a.ctypes.data if a.size else 0


def grid3(N):
    """N: an int (square / cubic grid) or a tuple (nx, ny[, nz])."""
    if isinstance(N, (int, np.integer)):
        return int(N), int(N), int(N)
    g = tuple(int(v) for v in N)
    return (g[0], g[1], 1) if len(g) == 2 else g


def stencil_rows(kind: int, N) -> int:
    return int(lib().hpcla_synth_stencil_rows(kind, *grid3(N)))


def stencil_local(kind: int, N, row_begin: int, row_end: int, T, Ti):
    """Rows [row_begin, row_end) (0-based) -> (rowptr 1-based, GLOBAL columns 1-based, values)."""
    L = lib()
    T, Ti = np.dtype(T), np.dtype(Ti)
    nnz = int(L.hpcla_synth_stencil_nnz(kind, *grid3(N), row_begin, row_end))
    if Ti == np.int32 and nnz >= 2**31 - 1:
        raise ValueError("local nnz does not fit Int32 row pointers")
    rowptr = np.empty(row_end - row_begin + 1, dtype=Ti)
    cols = np.empty(nnz, dtype=Ti)
    vals = np.empty(nnz, dtype=T)
    _check(L.hpcla_synth_stencil_fill(kind, *grid3(N), _DTYPE[T], _ITYPE[Ti], row_begin, row_end, _p(rowptr), _p(cols), _p(vals)))
    return rowptr, cols, vals


def powerlaw_local(n: int, seed: int, max_len: int, row_begin: int, row_end: int, T, Ti):
    L = lib()
    T, Ti = np.dtype(T), np.dtype(Ti)
    nnz = int(L.hpcla_synth_powerlaw_nnz(n, seed, max_len, row_begin, row_end))
    if Ti == np.int32 and nnz >= 2**31 - 1:
        raise ValueError("local nnz does not fit Int32 row pointers")
    rowptr = np.empty(row_end - row_begin + 1, dtype=Ti)
    cols = np.empty(nnz, dtype=Ti)
    vals = np.empty(nnz, dtype=T)
    _check(L.hpcla_synth_powerlaw_fill(n, seed, max_len, _DTYPE[T], _ITYPE[Ti], row_begin, row_end, _p(rowptr), _p(cols), _p(vals)))
    return rowptr, cols, vals


def vector_local(T, seed: int, begin: int, end: int) -> np.ndarray:
    """x[g] = 2u(g) - 1 for g in [begin, end) (0-based)."""
    out = np.empty(end - begin, dtype=np.dtype(T))
    _check(lib().hpcla_synth_vector(_DTYPE[np.dtype(T)], seed, begin, end, _p(out)))
    return out


def vector_at(T, seed: int, idx0) -> np.ndarray:
    """x at arbitrary 0-based global indices (runs of consecutive indices are generated in one call each)."""
    idx0 = np.asarray(idx0, dtype=np.int64)
    out = np.empty(idx0.shape, dtype=np.dtype(T))
    if idx0.size == 0:
        return out
    breaks = np.flatnonzero(np.diff(idx0) != 1) + 1
    starts = np.concatenate([[0], breaks])
    ends = np.concatenate([breaks, [idx0.size]])
    for s, e in zip(starts, ends):
        out[s:e] = vector_local(T, seed, int(idx0[s]), int(idx0[e - 1]) + 1)
    return out
