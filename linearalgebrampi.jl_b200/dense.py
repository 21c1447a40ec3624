"""HPCMatrix (distributed dense matrix, row-partitioned) — the part of src/dense.jl the sparse x dense product needs.

`HPCMatrix{T,B}` (src/dense.jl:59-69): structural_hash (lazy), row_partition, col_partition, A (this rank's rows, all
columns; Julia `Matrix{T}`, i.e. COLUMN-major), backend.  Here `A` is a 2-D torch CUDA tensor of shape
(local rows, ncols) whose memory is column-major (stride (1, ld)), so column k is a contiguous HPCVector slice exactly
as `B[:, k]` is in the reference (src/indexing.jl:385-393).  Everything else of dense.jl (dense x dense, transpose,
mapslices, indexing) is out of scope.

`A::HPCSparseMatrix * B::HPCMatrix` (src/sparse.jl:2391-2413) runs as ONE library call (hpcla_spmm_run): one halo
exchange for all columns and kernels that stage each tile of A once per 4 columns, instead of the reference's loop of
ncols SpMVs with a column extraction and an exchange each.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib
from .backends import HPCBackend, comm_allgather, comm_allreduce, comm_barrier, comm_rank, comm_size
from .vectors import HPCVector, _current_stream, _torch_dtype, compute_partition_hash, uniform_partition


def _colmajor_device(block: np.ndarray, backend: HPCBackend):
    """Host (rows, cols) array -> device tensor of the same shape with column-major memory."""
    import torch

    t = torch.from_numpy(np.ascontiguousarray(block.T)).to(backend.torch_device())  # (cols, rows) row-major
    return t.T  # (rows, cols) view, stride (1, rows)


class HPCMatrix:
    """HPCMatrix{T,B} (src/dense.jl:59-69)."""

    __slots__ = ("structural_hash", "row_partition", "col_partition", "A", "backend")

    def __init__(self, structural_hash, row_partition: np.ndarray, col_partition: np.ndarray, A, backend: HPCBackend):
        self.structural_hash = structural_hash
        self.row_partition = row_partition
        self.col_partition = col_partition
        self.A = A
        self.backend = backend

    # -- constructors ------------------------------------------------------------------------------------------
    @staticmethod
    def from_global(M, backend: HPCBackend, row_partition: Optional[np.ndarray] = None, col_partition: Optional[np.ndarray] = None) -> "HPCMatrix":
        """HPCMatrix(M, backend; row_partition, col_partition) — src/dense.jl:185-202."""
        M = np.asarray(M)
        P, r = comm_size(backend.comm), comm_rank(backend.comm)
        rp = uniform_partition(M.shape[0], P) if row_partition is None else np.ascontiguousarray(row_partition, dtype=np.int64)
        cp = uniform_partition(M.shape[1], P) if col_partition is None else np.ascontiguousarray(col_partition, dtype=np.int64)
        block = M[int(rp[r]) - 1 : int(rp[r + 1]) - 1, :].astype(backend.T, copy=False)
        A = _colmajor_device(block, backend) if backend.is_cuda else np.asfortranarray(block)
        return HPCMatrix(None, rp, cp, A, backend)

    @staticmethod
    def from_local(A_local, backend: HPCBackend, col_partition: Optional[np.ndarray] = None) -> "HPCMatrix":
        """HPCMatrix_local(A_local, backend; col_partition) — src/dense.jl:125-158: row partition from an Allgather of
        the local row counts; all ranks must hold the same number of columns (collective error otherwise)."""
        P = comm_size(backend.comm)
        nrows, ncols = int(A_local.shape[0]), int(A_local.shape[1])
        info = comm_allgather(backend.comm, (nrows, ncols))
        if any(c != info[0][1] for _, c in info):
            raise ValueError(f"HPCMatrix_local: All ranks must have the same number of columns. Got column counts: {[c for _, c in info]}")
        rp = np.concatenate([[1], 1 + np.cumsum(np.asarray([n for n, _ in info], dtype=np.int64))]).astype(np.int64)
        cp = uniform_partition(ncols, P) if col_partition is None else np.ascontiguousarray(col_partition, dtype=np.int64)
        if isinstance(A_local, np.ndarray):
            A = _colmajor_device(A_local.astype(backend.T, copy=False), backend) if backend.is_cuda else np.asfortranarray(A_local.astype(backend.T))
        else:
            A = A_local  # device storage; made column-major on first use
        return HPCMatrix(None, rp, cp, A, backend)

    # -- basics --------------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return (int(self.row_partition[-1]) - 1, int(self.col_partition[-1]) - 1)

    @property
    def local_rows(self) -> int:
        return int(self.A.shape[0])

    def column(self, k: int) -> HPCVector:
        """B[:, k] (0-based k here) — src/indexing.jl:385-393: the local part of column k as an HPCVector whose
        partition is B's row partition."""
        if k < 0 or k >= self.shape[1]:
            raise IndexError(f"HPCMatrix column index out of bounds: k={k}, ncols={self.shape[1]}")
        col = self.A[:, k]
        if not isinstance(col, np.ndarray) and not col.is_contiguous():
            col = col.contiguous()
        return HPCVector(compute_partition_hash(self.row_partition), self.row_partition, col, self.backend)

    def local_values(self) -> np.ndarray:
        return np.array(self.A) if isinstance(self.A, np.ndarray) else self.A.detach().cpu().numpy()

    def to_global(self) -> np.ndarray:
        """Matrix(A): all rows on every rank."""
        parts = comm_allgather(self.backend.comm, self.local_values())
        return np.concatenate(parts, axis=0) if parts else np.zeros((0, self.shape[1]), dtype=self.backend.T)

    def __repr__(self):
        return f"HPCMatrix({self.shape[0]}x{self.shape[1]}, local rows={self.local_rows}, T={self.backend.T}, {self.backend.device})"


def _colmajor(t):
    """A (rows, cols) device tensor with column-major memory (copy only if it is not already)."""
    if t.dim() != 2:
        raise ValueError("HPCMatrix.A must be 2-D")
    if t.shape[1] <= 1 or (t.stride(0) == 1 and t.stride(1) >= max(t.shape[0], 1)):
        return t if t.stride(0) == 1 or t.shape[0] <= 1 else t.contiguous()
    return t.T.contiguous().T


def spmm(A, B: HPCMatrix) -> HPCMatrix:
    """Base.:*(A::HPCSparseMatrix, B::HPCMatrix) — src/sparse.jl:2391-2413.  Result: HPCMatrix with A's row partition
    and the uniform column partition HPCMatrix_local gives it (src/dense.jl:125-126)."""
    import torch

    from . import sparse as sp

    b = A.backend
    if not b.is_cuda or not B.backend.is_cuda:
        raise _lib.HPCLAError("A*B needs DeviceCUDA operands: this build has no CPU arithmetic (and no CPU fallback)")
    if B.backend.T != b.T:
        raise TypeError(f"element types differ: A is {b.T}, B is {B.backend.T}")
    if B.shape[0] != A.shape[1]:
        raise ValueError(f"DimensionMismatch: A has {A.shape[1]} columns, B has {B.shape[0]} rows")
    ncols = B.shape[1]
    P = comm_size(b.comm)
    col0 = B.column(0) if ncols > 0 else HPCVector.zeros(b, B.shape[0], partition=B.row_partition)
    plan = sp.get_vector_plan(A, col0)  # one plan for every column: they share B's row partition
    op = sp._bound_op(A, plan, col0)
    Bl = _colmajor(B.A)
    C = torch.empty((ncols, A.nrows_local), dtype=_torch_dtype(b.T), device=b.torch_device()).T  # column-major (rows, cols)
    ldb = Bl.stride(1) if ncols > 1 else max(Bl.shape[0], 1)
    ldc = max(A.nrows_local, 1)
    L = _lib.lib()
    # every rank must take the same route (the product is collective): decided once per plan, like the plan itself
    all_in_place = getattr(plan, "_all_ranks_in_place", None)
    if all_in_place is None:
        all_in_place = bool(comm_allreduce(b.comm, int(sp.spmv_info(A, col0)["x_in_place"]), "min"))
        try:
            plan._all_ranks_in_place = all_in_place
        except AttributeError:
            pass
    if ncols > 0 and not all_in_place:
        # own columns of A with holes in B's local rows: column by column, exactly the reference's loop
        for k in range(ncols):
            yk = sp.matvec(A, B.column(k))
            C[:, k] = yk.v
    elif ncols > 0:
        stream = _current_stream(b)
        if b.ctx().world == "threads":
            _lib.check(L.hpcla_spmm_begin(op, _lib.ptr(Bl), ldb, _lib.ptr(C), ldc, ncols, stream))
            comm_barrier(b.comm)
            _lib.check(L.hpcla_spmm_finish(op))
            comm_barrier(b.comm)
        else:
            _lib.check(L.hpcla_spmm_run(op, _lib.ptr(Bl), ldb, _lib.ptr(C), ldc, ncols, stream))
        C._hpcla_keepalive = Bl  # the enqueued kernels read Bl
    return HPCMatrix(None, A.row_partition.copy(), uniform_partition(ncols, P), C, b)
