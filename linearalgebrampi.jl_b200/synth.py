"""Deterministic synthetic inputs as HPCSparseMatrix / HPCVector objects.

The generators themselves live in their own host-only library (hpcla_synth/, include/hpcla_synth.h) shared with the
CPU oracle and the CPU reference arm, so that the GPU path and the checker see identical bits.  Each rank generates
only its own rows (HPCSparseMatrix_local route, src/sparse.jl:454): nothing global is ever materialised.
"""
from __future__ import annotations

import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import hpcla_synth as _gen  # noqa: E402
from hpcla_synth import (  # noqa: E402,F401
    LAPLACE2D_5PT, POISSON3D_7PT, POWERLAW_MAX_LEN, POWERLAW_SEED, STENCIL3D_27PT, X_SEED, powerlaw_local, stencil_local, stencil_rows,
    vector_at, vector_local,
)

from .backends import HPCBackend, comm_rank, comm_size  # noqa: E402
from .sparse import HPCSparseMatrix  # noqa: E402
from .vectors import HPCVector, uniform_partition  # noqa: E402


def _my_rows(n: int, backend: HPCBackend):
    part = uniform_partition(n, comm_size(backend.comm))
    r = comm_rank(backend.comm)
    return part, int(part[r]) - 1, int(part[r + 1]) - 1


def stencil_matrix(kind: int, N, backend: HPCBackend) -> HPCSparseMatrix:
    n = stencil_rows(kind, N)
    part, b, e = _my_rows(n, backend)
    rowptr, cols, vals = stencil_local(kind, N, b, e, backend.T, backend.Ti)
    return HPCSparseMatrix.from_local(rowptr, cols, vals, n, backend, col_partition=part)


def powerlaw_matrix(n: int, backend: HPCBackend, seed: int = POWERLAW_SEED, max_len: int = POWERLAW_MAX_LEN) -> HPCSparseMatrix:
    part, b, e = _my_rows(n, backend)
    rowptr, cols, vals = powerlaw_local(n, seed, max_len, b, e, backend.T, backend.Ti)
    return HPCSparseMatrix.from_local(rowptr, cols, vals, n, backend, col_partition=part)


def vector(n: int, backend: HPCBackend, seed: int = X_SEED) -> HPCVector:
    part, b, e = _my_rows(n, backend)
    return HPCVector.from_local(vector_local(backend.T, seed, b, e), backend)
