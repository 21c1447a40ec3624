"""Deterministic synthetic inputs — python front-end of include/hpcla_synth.h.

Used by tests, bench.py and the CPU baseline alike, so that the GPU path and the oracle see identical bits.  Each rank
generates only its own rows (HPCSparseMatrix_local route, src/sparse.jl:454): nothing global is ever materialised.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .backends import HPCBackend, comm_rank, comm_size
from .sparse import HPCSparseMatrix
from .vectors import HPCVector, uniform_partition

LAPLACE2D_5PT, POISSON3D_7PT, STENCIL3D_27PT = 0, 1, 2
X_SEED = 0x5EED


def _grid(N):
    """N: an int (square / cubic grid) or a tuple (nx, ny[, nz])."""
    if isinstance(N, (int, np.integer)):
        return int(N), int(N), int(N)
    g = tuple(int(v) for v in N)
    return (g[0], g[1], 1) if len(g) == 2 else g


def stencil_rows(kind: int, N) -> int:
    return int(_lib.lib().hpcla_synth_stencil_rows(kind, *_grid(N)))


def stencil_local(kind: int, N, row_begin: int, row_end: int, T, Ti):
    """Rows [row_begin, row_end) (0-based) -> (rowptr 1-based, GLOBAL columns 1-based, values)."""
    L = _lib.lib()
    T, Ti = np.dtype(T), np.dtype(Ti)
    nnz = int(L.hpcla_synth_stencil_nnz(kind, *_grid(N), row_begin, row_end))
    if Ti == np.int32 and nnz >= 2**31 - 1:
        raise _lib.HPCLAError("local nnz does not fit Int32 row pointers")
    rowptr = np.empty(row_end - row_begin + 1, dtype=Ti)
    cols = np.empty(nnz, dtype=Ti)
    vals = np.empty(nnz, dtype=T)
    _lib.check(L.hpcla_synth_stencil_fill(kind, *_grid(N), _lib.dtype_code(T), _lib.itype_code(Ti), row_begin, row_end, _lib.ptr(rowptr), _lib.ptr(cols), _lib.ptr(vals)))
    return rowptr, cols, vals


def powerlaw_local(n: int, seed: int, max_len: int, row_begin: int, row_end: int, T, Ti):
    L = _lib.lib()
    T, Ti = np.dtype(T), np.dtype(Ti)
    nnz = int(L.hpcla_synth_powerlaw_nnz(n, seed, max_len, row_begin, row_end))
    if Ti == np.int32 and nnz >= 2**31 - 1:
        raise _lib.HPCLAError("local nnz does not fit Int32 row pointers")
    rowptr = np.empty(row_end - row_begin + 1, dtype=Ti)
    cols = np.empty(nnz, dtype=Ti)
    vals = np.empty(nnz, dtype=T)
    _lib.check(L.hpcla_synth_powerlaw_fill(n, seed, max_len, _lib.dtype_code(T), _lib.itype_code(Ti), row_begin, row_end, _lib.ptr(rowptr), _lib.ptr(cols), _lib.ptr(vals)))
    return rowptr, cols, vals


def vector_local(T, seed: int, begin: int, end: int) -> np.ndarray:
    """x[g] = 2u(g) - 1 for g in [begin, end) (0-based)."""
    out = np.empty(end - begin, dtype=np.dtype(T))
    _lib.check(_lib.lib().hpcla_synth_vector(_lib.dtype_code(T), seed, begin, end, _lib.ptr(out)))
    return out


def _my_rows(n: int, backend: HPCBackend):
    part = uniform_partition(n, comm_size(backend.comm))
    r = comm_rank(backend.comm)
    return part, int(part[r]) - 1, int(part[r + 1]) - 1


def stencil_matrix(kind: int, N, backend: HPCBackend) -> HPCSparseMatrix:
    n = stencil_rows(kind, N)
    part, b, e = _my_rows(n, backend)
    rowptr, cols, vals = stencil_local(kind, N, b, e, backend.T, backend.Ti)
    return HPCSparseMatrix.from_local(rowptr, cols, vals, n, backend, col_partition=part)


def powerlaw_matrix(n: int, backend: HPCBackend, seed: int = 0xC4, max_len: int = 1_000_000) -> HPCSparseMatrix:
    part, b, e = _my_rows(n, backend)
    rowptr, cols, vals = powerlaw_local(n, seed, max_len, b, e, backend.T, backend.Ti)
    return HPCSparseMatrix.from_local(rowptr, cols, vals, n, backend, col_partition=part)


def vector(n: int, backend: HPCBackend, seed: int = X_SEED) -> HPCVector:
    part, b, e = _my_rows(n, backend)
    return HPCVector.from_local(vector_local(backend.T, seed, b, e), backend)
