"""HPCSparseMatrix, VectorPlan and the multiply operators — host-side mirror of src/sparse.jl for the SpMV hot path.

Mirrors, with the reference's names and semantics:
  * `HPCSparseMatrix{T,Ti,B}` and its fields                     (src/sparse.jl:319-337)
  * `HPCSparseMatrix(A_global, backend; row_partition, col_partition)`  (:398-413)
  * `HPCSparseMatrix_local(A_local, backend; col_partition)`     (:454-525)
  * `compute_structural_hash` / `_ensure_hash`                   (:97-121; src/HPCLinearAlgebra.jl:759-764)
  * `VectorPlan(A, x)`, `get_vector_plan`, `_vector_plan_cache`  (:1875-2001; src/HPCLinearAlgebra.jl:133)
  * `execute_plan!(plan, x)`                                     (src/vectors.jl:394-463)
  * `mul!(y, A, x)`, `A * x`, `transpose(A) * x`, `transpose(v) * A`   (:2019-2037, 2096-2128, 2375-2379, 2136-2142)
  * `TransposePlan` + cached materialisation                     (:1551-1865)
  * `clear_plan_cache!`, `cache_sizes`, `to_backend`             (src/HPCLinearAlgebra.jl:181-244, 337-378)

The integer logic runs in libhpcla_b200.so's host functions, the per-call path (halo exchange + kernels) in its CUDA
part; this file only moves pointers.  Communication for plan construction is done here by the comm primitives of
backends.py (the reference does it with MPI at the same places).
"""
from __future__ import annotations

import ctypes
import os
import itertools
import weakref
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from .backends import (
    CommThreads,
    HPCBackend,
    assert_backends_compatible,
    comm_allgather,
    comm_barrier,
    comm_exchange,
    comm_rank,
    comm_size,
    comm_uid,
)
from .vectors import HPCVector, _current_stream, _digest, _to_device, _torch_dtype, compute_partition_hash, uniform_partition

# module-level plan cache (src/HPCLinearAlgebra.jl:133) — never evicts; clear_plan_cache() wipes it
_vector_plan_cache: Dict[tuple, "VectorPlan"] = {}
# counts plan constructions (test hook for the memoisation behaviour)
plan_build_count = 0
# matrices that may hold bound operators: clear_plan_cache() drops those operators with the plans they were bound to
_live_matrices: "weakref.WeakSet[HPCSparseMatrix]" = weakref.WeakSet()


class _DevView:
    """Zero-copy torch view of library-owned device memory (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int, np_dtype, keep=None):
        np_dtype = np.dtype(np_dtype)
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": np_dtype.str, "data": (ptr, False), "version": 3, "strides": None}
        self._keep = keep


# ---------------------------------------------------------------------------------------------------------------
# HPCSparseMatrix
# ---------------------------------------------------------------------------------------------------------------
class HPCSparseMatrix:
    """HPCSparseMatrix{T,Ti,B} (src/sparse.jl:319-337).  All index arrays hold the reference's 1-based values."""

    def __init__(self, structural_hash, row_partition, col_partition, col_indices, rowptr, colval, nzval, nrows_local,
                 ncols_compressed, rowptr_target, colval_target, backend: HPCBackend):
        self.structural_hash: Optional[bytes] = structural_hash
        self.row_partition: np.ndarray = row_partition
        self.col_partition: np.ndarray = col_partition
        self.col_indices: np.ndarray = col_indices  # Int64, sorted global columns present locally
        self.rowptr: np.ndarray = rowptr  # host, Ti
        self.colval: np.ndarray = colval  # host, Ti, LOCAL index into col_indices
        self.nzval = nzval  # device storage (torch CUDA tensor) or numpy on DeviceCPU
        self.nrows_local: int = nrows_local
        self.ncols_compressed: int = ncols_compressed
        self.cached_transpose: Optional["HPCSparseMatrix"] = None
        self.cached_symmetric: Optional[bool] = None
        self.rowptr_target = rowptr_target  # device copies of the structure (src/sparse.jl:334-335)
        self.colval_target = colval_target
        self.backend = backend
        # backend-derived device state lives beside the reference's fields (SURVEY §8b "side cache")
        self._csr: Optional[int] = None
        self._ops: Dict[tuple, int] = {}
        self._graphs: Dict[tuple, tuple] = {}
        _live_matrices.add(self)

    # -- constructors ------------------------------------------------------------------------------------------
    @staticmethod
    def from_global(A, backend: HPCBackend, row_partition=None, col_partition=None) -> "HPCSparseMatrix":
        """HPCSparseMatrix{T}(A::SparseMatrixCSC, backend; row_partition, col_partition) — src/sparse.jl:398-413.
        `A` is a scipy.sparse matrix, identical on all ranks (canonicalised like Julia's `sparse`: duplicates summed,
        explicit zeros kept, ascending columns within a row)."""
        import scipy.sparse as sp

        A = sp.csr_matrix(A)
        A.sum_duplicates()
        A.sort_indices()
        m, n = A.shape
        P, r = comm_size(backend.comm), comm_rank(backend.comm)
        rp = uniform_partition(m, P) if row_partition is None else np.ascontiguousarray(row_partition, dtype=np.int64)
        cp = uniform_partition(n, P) if col_partition is None else np.ascontiguousarray(col_partition, dtype=np.int64)
        r0, r1 = int(rp[r]) - 1, int(rp[r + 1]) - 1  # local row range (:405-409)
        lo, hi = int(A.indptr[r0]), int(A.indptr[r1])
        rowptr1 = (A.indptr[r0 : r1 + 1].astype(np.int64) - lo + 1).astype(backend.Ti)
        gcols1 = (A.indices[lo:hi].astype(np.int64) + 1).astype(backend.Ti)
        return HPCSparseMatrix.from_local(rowptr1, gcols1, A.data[lo:hi].astype(backend.T), n, backend, col_partition=cp)

    @staticmethod
    def from_local(rowptr, global_cols, nzval, ncols_global: int, backend: HPCBackend, col_partition=None) -> "HPCSparseMatrix":
        """HPCSparseMatrix_local(A_local::SparseMatrixCSR, backend; col_partition) — src/sparse.jl:454-525.
        rowptr: 1-based [nrows_local+1]; global_cols: 1-based GLOBAL columns, ascending within a row; nzval: values."""
        Ti, T = backend.Ti, backend.T
        comm = backend.comm
        P = comm_size(comm)
        rowptr = np.ascontiguousarray(rowptr, dtype=Ti)  # :466-474 index arrays converted to the backend's Ti
        global_cols = np.ascontiguousarray(global_cols, dtype=Ti)
        nzval_cpu = np.ascontiguousarray(nzval, dtype=T)
        local_nrows = len(rowptr) - 1
        nnz = int(rowptr[-1]) - 1
        if nnz != len(global_cols) or nnz != len(nzval_cpu):
            raise ValueError("HPCSparseMatrix_local: rowptr, column and value arrays disagree on nnz")
        # :477-497 Allgather [local_nrows, ncols_global] -> row_partition; all ranks must agree on the column count
        info = comm_allgather(comm, (local_nrows, int(ncols_global)))
        if any(c != info[0][1] for _, c in info):
            raise ValueError(f"HPCSparseMatrix_local: All ranks must have the same number of columns. Got column counts: {[c for _, c in info]}")
        row_partition = np.concatenate([[1], 1 + np.cumsum(np.array([c for c, _ in info], dtype=np.int64))]).astype(np.int64)
        cp = uniform_partition(ncols_global, P) if col_partition is None else np.ascontiguousarray(col_partition, dtype=np.int64)
        # :501-504 col_indices = unique!(sort(cols)); colval = searchsortedfirst(col_indices, col)
        colval = np.empty(nnz, dtype=Ti)
        col_indices = np.empty(min(nnz, int(ncols_global)), dtype=np.int64)
        ncc = ctypes.c_int64(0)
        _lib.check(_lib.lib().hpcla_compress_columns(_lib.itype_code(Ti), nnz, _lib.ptr(global_cols), int(ncols_global), _lib.ptr(colval), _lib.ptr(col_indices), ctypes.byref(ncc)))
        col_indices = col_indices[: ncc.value].copy()
        # :517-520 device copies
        return HPCSparseMatrix(None, row_partition, cp, col_indices, rowptr, colval, _to_device(nzval_cpu, backend), local_nrows,
                               int(ncc.value), _to_device(rowptr, backend), _to_device(colval, backend), backend)

    # -- basics --------------------------------------------------------------------------------------------------
    @property
    def shape(self) -> Tuple[int, int]:
        return int(self.row_partition[-1]) - 1, int(self.col_partition[-1]) - 1

    @property
    def nnz_local(self) -> int:
        return int(self.rowptr[-1]) - 1

    def nzval_host(self) -> np.ndarray:
        return self.nzval if isinstance(self.nzval, np.ndarray) else self.nzval.detach().cpu().numpy()

    def __matmul__(self, x):
        return self.__mul__(x)

    def __mul__(self, x):
        if isinstance(x, HPCVector):
            return matvec(self, x)
        from .dense import HPCMatrix, spmm

        if isinstance(x, HPCMatrix):  # Base.:*(A::HPCSparseMatrix, B::HPCMatrix) — src/sparse.jl:2391-2413
            return spmm(self, x)
        if isinstance(x, HPCSparseMatrix):  # Base.:*(A::HPCSparseMatrix, B::HPCSparseMatrix) — src/sparse.jl:991-1059
            from .spgemm import spgemm

            return spgemm(self, x)
        return NotImplemented

    @property
    def T(self) -> "Transpose":
        return Transpose(self)

    def invalidate_structure(self) -> None:
        """What a structural setindex! does (src/indexing.jl:1291-1294): drop the hash and the cached transpose.
        Device-side derived state is keyed the same way and dropped with them."""
        self.structural_hash = None
        _drop_device_state(self)
        if self.cached_transpose is not None:
            other, self.cached_transpose = self.cached_transpose, None
            other.cached_transpose = None

    def values_changed(self) -> None:
        """What an in-place value setindex! does (src/indexing.jl:978-979): keep the hash, drop the cached transpose.
        The kernels read A.nzval itself, so nothing else needs refreshing."""
        if self.cached_transpose is not None:
            other, self.cached_transpose = self.cached_transpose, None
            other.cached_transpose = None

    def __del__(self):
        try:
            _drop_device_state(self)
        except Exception:
            pass

    def __repr__(self):
        m, n = self.shape
        return f"HPCSparseMatrix({m}x{n}, local rows={self.nrows_local}, nnz_local={self.nnz_local}, T={self.backend.T}, Ti={self.backend.Ti}, {self.backend.device})"


def _drop_bound_ops(A: HPCSparseMatrix) -> None:
    L = _lib.lib()
    for op in A._ops.values():
        L.hpcla_spmv_destroy(op)
    A._ops.clear()
    A._graphs.clear()


def _drop_device_state(A: HPCSparseMatrix) -> None:
    L = _lib.lib()
    _drop_bound_ops(A)
    if A._csr is not None:
        L.hpcla_csr_destroy(A._csr)
        A._csr = None


def compute_structural_hash(row_partition, col_indices, rowptr, colval, comm) -> bytes:
    """src/sparse.jl:97-121: Blake3 over length-prefixed (row_partition, col_indices, rowptr, colval), Allgathered and
    re-hashed so that every rank holds the same key.  A cache key, not a numerical result."""
    chunks = []
    for a in (row_partition, col_indices, rowptr, colval):
        a = np.ascontiguousarray(a)
        chunks.append(np.int64(a.size).tobytes())
        chunks.append(a.tobytes() if a.size < (1 << 20) else memoryview(a).cast("B"))
    local = _digest(*chunks)
    return _digest(*comm_allgather(comm, local))


def _ensure_hash(A: HPCSparseMatrix) -> bytes:
    """src/HPCLinearAlgebra.jl:759-764 (lazy, collective on first use)."""
    if A.structural_hash is None:
        A.structural_hash = compute_structural_hash(A.row_partition, A.col_indices, A.rowptr, A.colval, A.backend.comm)
    return A.structural_hash


# ---------------------------------------------------------------------------------------------------------------
# VectorPlan
# ---------------------------------------------------------------------------------------------------------------
class VectorPlan:
    """VectorPlan{T,Ti,AV} (src/vectors.jl:229-251): the index fields, exported from the library's plan object."""

    _uids = itertools.count(1)

    def __init__(self, handle: int, Ti, n_x_local: int):
        L = _lib.lib()
        self.handle = handle
        self.uid = next(VectorPlan._uids)
        self.n_x_local = n_x_local
        Ti = np.dtype(Ti)

        def get(field, slot=0):
            n = ctypes.c_int64()
            _lib.check(L.hpcla_plan_len(handle, field, slot, ctypes.byref(n)))
            a = np.empty(n.value, dtype=np.int64)
            _lib.check(L.hpcla_plan_get(handle, field, slot, _lib.ptr(a)))
            return a

        self.send_rank_ids = get(0)
        self.recv_rank_ids = get(1)
        self.local_src_indices = get(2).astype(Ti)
        self.local_dst_indices = get(3).astype(Ti)
        self.send_indices: List[np.ndarray] = [get(4, i).astype(Ti) for i in range(len(self.send_rank_ids))]
        self.recv_perm: List[np.ndarray] = [get(5, i).astype(Ti) for i in range(len(self.recv_rank_ids))]
        n = ctypes.c_int64()
        _lib.check(L.hpcla_plan_n_gathered(handle, ctypes.byref(n)))
        self.n_gathered = n.value
        self.result_partition_hash: Optional[bytes] = None  # cached lazily by A*x (src/sparse.jl:2103-2106)
        self.result_partition: Optional[np.ndarray] = None

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().hpcla_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def build_vector_plan(A: HPCSparseMatrix, x: HPCVector) -> VectorPlan:
    """VectorPlan(A, x) — src/sparse.jl:1875-1984.  Collective."""
    global plan_build_count
    assert_backends_compatible(A.backend, x.backend)  # :1876
    L = _lib.lib()
    comm = A.backend.comm
    rank, P = comm_rank(comm), comm_size(comm)
    ci = np.ascontiguousarray(A.col_indices, dtype=np.int64)
    xp = np.ascontiguousarray(x.partition, dtype=np.int64)
    if len(xp) != P + 1:
        raise ValueError("VectorPlan: x.partition does not describe this communicator")
    pb = ctypes.c_void_p()
    _lib.check(L.hpcla_plan_begin(rank, P, _lib.ptr(ci), len(ci), _lib.ptr(xp), ctypes.byref(pb)))
    counts = np.zeros(P, dtype=np.int64)
    _lib.check(L.hpcla_planb_counts(pb, _lib.ptr(counts)))
    send = {}
    for o in range(P):  # Step 3 (:1908-1918): the global indices wanted from each owner
        if o != rank and counts[o] > 0:
            buf = np.empty(int(counts[o]), dtype=np.int64)
            _lib.check(L.hpcla_planb_requests(pb, o, _lib.ptr(buf)))
            send[o] = buf
    got = comm_exchange(comm, send, np.int64, tag=20)  # Alltoall + tag-20 Isend/Irecv (:1899-1936)
    recv_counts = np.zeros(P, dtype=np.int64)
    lists = [None] * P
    for q, a in got.items():
        recv_counts[q] = len(a)
        lists[q] = np.ascontiguousarray(a, dtype=np.int64)
    ph = ctypes.c_void_p()
    _lib.check(L.hpcla_plan_finish(pb, _lib.ptr(recv_counts), _lib.ptr_array(lists), ctypes.byref(ph)))
    plan_build_count += 1
    return VectorPlan(ph.value, A.backend.Ti, x.local_size)


def import_vector_plan(A: HPCSparseMatrix, x: HPCVector, send_rank_ids, send_indices, recv_rank_ids, recv_perm, local_src_indices,
                       local_dst_indices, n_gathered: int) -> VectorPlan:
    """Adopt a VectorPlan that was built elsewhere — the route of the Julia binding, which hands the plan the reference
    itself built (src/sparse.jl:1875-1984; fields of src/vectors.jl:229-251, index arrays in Ti width, 1-based) to
    hpcla_plan_import — and memoise it under the reference's cache key, so that `A * x` uses it."""
    L = _lib.lib()
    Ti = np.dtype(A.backend.Ti)
    comm = A.backend.comm
    sids = np.ascontiguousarray(send_rank_ids, dtype=np.int64)
    rids = np.ascontiguousarray(recv_rank_ids, dtype=np.int64)
    sidx = [np.ascontiguousarray(a, dtype=Ti) for a in send_indices]
    rprm = [np.ascontiguousarray(a, dtype=Ti) for a in recv_perm]
    slen = np.array([len(a) for a in sidx], dtype=np.int64)
    rlen = np.array([len(a) for a in rprm], dtype=np.int64)
    lsrc = np.ascontiguousarray(local_src_indices, dtype=Ti)
    ldst = np.ascontiguousarray(local_dst_indices, dtype=Ti)
    ph = ctypes.c_void_p()
    _lib.check(L.hpcla_plan_import(comm_rank(comm), comm_size(comm), _lib.itype_code(Ti), int(n_gathered), x.local_size, len(sids), _lib.ptr(sids),
                                   _lib.ptr(slen), _lib.ptr_array(sidx), len(rids), _lib.ptr(rids), _lib.ptr(rlen), _lib.ptr_array(rprm), len(lsrc),
                                   _lib.ptr(lsrc), _lib.ptr(ldst), ctypes.byref(ph)))
    plan = VectorPlan(ph.value, Ti, x.local_size)
    key = (_ensure_hash(A), x.structural_hash, A.backend.T.str, A.backend.Ti.str, _storage_tag(x), _comm_key(A.backend))
    _vector_plan_cache[key] = plan
    return plan


def _storage_tag(x: HPCVector) -> str:
    return "Vector" if isinstance(x.v, np.ndarray) else "CuVector"


def get_vector_plan(A: HPCSparseMatrix, x: HPCVector) -> VectorPlan:
    """get_vector_plan — src/sparse.jl:1992-2001: memoised on (hash(A), hash(x.partition), T, Ti, typeof(x.v))."""
    key = (_ensure_hash(A), x.structural_hash, A.backend.T.str, A.backend.Ti.str, _storage_tag(x), _comm_key(A.backend))
    plan = _vector_plan_cache.get(key)
    if plan is None:
        plan = build_vector_plan(A, x)
        _vector_plan_cache[key] = plan
    return plan


def _comm_key(b: HPCBackend):
    # one python process may host several ranks (CommThreads): their plans differ, so the rank is part of the key
    return (comm_uid(b.comm), comm_rank(b.comm))


def clear_plan_cache() -> None:
    """clear_plan_cache!() — src/HPCLinearAlgebra.jl:181-201."""
    from . import vectors as _v

    from .spgemm import _matrix_plan_cache  # (the package attribute `spgemm` is the function, not the module)

    _vector_plan_cache.clear()
    _v._repartition_plan_cache.clear()
    _matrix_plan_cache.clear()
    # operators are bound to plans: with the plans gone they would only be re-created under new plan ids, so they
    # (gathered / send buffers, tile lists, compact tile data) are released here instead of living as long as A does
    for A in list(_live_matrices):
        _drop_bound_ops(A)


def cache_sizes() -> Dict[str, int]:
    """cache_sizes() — src/HPCLinearAlgebra.jl:208-244 (only the cache this path owns)."""
    from . import vectors as _v

    from .spgemm import _matrix_plan_cache

    return {"vector_plan": len(_vector_plan_cache), "repartition_plan": len(_v._repartition_plan_cache), "matrix_plan": len(_matrix_plan_cache)}


# ---------------------------------------------------------------------------------------------------------------
# device binding
# ---------------------------------------------------------------------------------------------------------------
def _csr_handle(A: HPCSparseMatrix) -> int:
    if A._csr is None:
        b = A.backend
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().hpcla_csr_create(b.ctx().handle, _lib.dtype_code(b.T), _lib.itype_code(b.Ti), A.nrows_local, A.ncols_compressed,
                                               A.nnz_local, _lib.ptr(A.rowptr_target), _lib.ptr(A.colval_target), _lib.ptr(A.nzval), ctypes.byref(h)))
        A._csr = h.value
    return A._csr


def _bound_op(A: HPCSparseMatrix, plan: VectorPlan, x: HPCVector) -> int:
    key = (plan.uid, x.local_size)
    op = A._ops.get(key)
    if op is None:
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().hpcla_spmv_create(A.backend.ctx().handle, _csr_handle(A), plan.handle, x.local_size, ctypes.byref(h)))
        op = A._ops[key] = h.value
    return op


def spmv_info(A: HPCSparseMatrix, x: HPCVector) -> Dict[str, int]:
    """Introspection: tiles, interior/boundary split, whether x.v is read in place (tests)."""
    L = _lib.lib()
    op = _bound_op(A, get_vector_plan(A, x), x)
    ni, nb = ctypes.c_int64(), ctypes.c_int64()
    xin, sc = ctypes.c_int(), ctypes.c_int()
    _lib.check(L.hpcla_spmv_info(op, ctypes.byref(ni), ctypes.byref(nb), ctypes.byref(xin), ctypes.byref(sc)))
    nt, nl, var = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    _lib.check(L.hpcla_csr_info(_csr_handle(A), ctypes.byref(nt), ctypes.byref(nl), ctypes.byref(var)))
    nr, ng, ne, win = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    _lib.check(L.hpcla_csr_tile_classes(_csr_handle(A), ctypes.byref(nr), ctypes.byref(ng), ctypes.byref(ne), ctypes.byref(win)))
    lists = (ctypes.c_int64 * 6)()
    _lib.check(L.hpcla_spmv_tile_lists(op, lists))
    return {"tiles": nt.value, "long_rows": nl.value, "lanes_per_row": var.value, "rowwalk_tiles": nr.value, "general_tiles": ng.value,
            "empty_tiles": ne.value, "tile_window": win.value, "interior_tiles": ni.value, "boundary_tiles": nb.value,
            "x_in_place": xin.value, "sends_contiguous": sc.value, "launches": int(L.hpcla_spmv_launch_count(op)),
            "compact_tiles": int(lists[4]), "flat_chunks": int(lists[5]), "plain_interior_rowwalk_tiles": int(lists[0])}


def _check_mul_args(A: HPCSparseMatrix, x: HPCVector):
    if not A.backend.is_cuda or not x.backend.is_cuda:
        raise _lib.HPCLAError("A*x needs DeviceCUDA operands: this build has no CPU arithmetic (and no CPU fallback)")
    if x.backend.T != A.backend.T:
        raise TypeError(f"element types differ: A is {A.backend.T}, x is {x.backend.T}")
    if len(x) != A.shape[1]:
        raise ValueError(f"DimensionMismatch: A has {A.shape[1]} columns, x has length {len(x)}")


def _run_multiply(A: HPCSparseMatrix, x: HPCVector, y_local) -> None:
    L = _lib.lib()
    plan = get_vector_plan(A, x)
    op = _bound_op(A, plan, x)
    stream = _current_stream(A.backend)
    if A.backend.ctx().world == "threads":
        _lib.check(L.hpcla_spmv_begin(op, _lib.ptr(x.v), _lib.ptr(y_local), stream))
        comm_barrier(A.backend.comm)
        _lib.check(L.hpcla_spmv_finish(op))
        comm_barrier(A.backend.comm)
    else:
        _lib.check(L.hpcla_spmv_run(op, _lib.ptr(x.v), _lib.ptr(y_local), stream))


def execute_plan(plan: VectorPlan, A: HPCSparseMatrix, x: HPCVector):
    """execute_plan!(plan, x) — src/vectors.jl:394-463: returns `gathered` (device view, gathered[d] ==
    x_global[col_indices[d]]).  The multiply itself never materialises the own segment; this is the parity hook."""
    import torch

    L = _lib.lib()
    op = _bound_op(A, plan, x)
    out = ctypes.c_void_p()
    _lib.check(L.hpcla_spmv_gather(op, _lib.ptr(x.v), _current_stream(A.backend), ctypes.byref(out)))
    if A.backend.ctx().world == "threads":
        comm_barrier(A.backend.comm)
        _lib.check(L.hpcla_spmv_gather_finish(op))
        comm_barrier(A.backend.comm)
    if plan.n_gathered == 0:
        return torch.zeros(0, dtype=_torch_dtype(A.backend.T), device=A.backend.torch_device())
    view_dtype = np.float64 if A.backend.T == np.complex128 else A.backend.T
    n = plan.n_gathered * (2 if A.backend.T == np.complex128 else 1)
    t = torch.as_tensor(_DevView(out.value, n, view_dtype, keep=A), device=A.backend.torch_device())
    if A.backend.T == np.complex128:
        t = torch.view_as_complex(t.view(-1, 2))
    return t


def mul(y: HPCVector, A: HPCSparseMatrix, x: HPCVector) -> HPCVector:
    """LinearAlgebra.mul!(y, A, x) — src/sparse.jl:2019-2037.  In place on y.v; returns y."""
    _check_mul_args(A, x)
    if y.local_size != A.nrows_local:
        raise ValueError(f"DimensionMismatch: y holds {y.local_size} local rows, A has {A.nrows_local}")
    _run_multiply(A, x, y.v)
    return y


class HostBuffer:
    """Pinned host memory next to this rank's GPU (hpcla_host_alloc): `.array` is a numpy view of it.  For the host
    sides of mul_staged; on a multi-socket box the copies then stay on the GPU's own NUMA node."""

    def __init__(self, backend: HPCBackend, n: int, dtype=None):
        dt = np.dtype(backend.T if dtype is None else dtype)
        self._ctx = backend.ctx()
        p = ctypes.c_void_p()
        node = ctypes.c_int(-1)
        nbytes = int(n) * dt.itemsize
        _lib.check(_lib.lib().hpcla_host_alloc(self._ctx.handle, nbytes, ctypes.byref(p), ctypes.byref(node)))
        self.ptr = p.value
        self.numa_node = node.value
        raw = (ctypes.c_byte * max(nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(raw, dtype=dt, count=int(n))

    def close(self):
        if self.ptr:
            self.array = None
            _lib.lib().hpcla_host_free(self._ctx.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def host_buffer(backend: HPCBackend, n: int, dtype=None) -> HostBuffer:
    return HostBuffer(backend, n, dtype)


def mul_graph(y: HPCVector, A: HPCSparseMatrix, x: HPCVector) -> HPCVector:
    """mul!(y, A, x) replayed from a CUDA graph bound to these x.v / y.v (hpcla_spmv_graph_capture / _launch): the first
    call with a given (x.v, y.v) runs one plain multiply (so every lazily configured kernel attribute is set) and then
    captures; later calls are one graph launch.  NCCL world or a single rank.  Collective."""
    _check_mul_args(A, x)
    if y.local_size != A.nrows_local:
        raise ValueError(f"DimensionMismatch: y holds {y.local_size} local rows, A has {A.nrows_local}")
    L = _lib.lib()
    op = _bound_op(A, get_vector_plan(A, x), x)
    stream = _current_stream(A.backend)
    key = (op, _lib.ptr(x.v), _lib.ptr(y.v))
    if A._graphs.get("bound") != key:
        _lib.check(L.hpcla_spmv_run(op, _lib.ptr(x.v), _lib.ptr(y.v), stream))
        _lib.check(L.hpcla_spmv_graph_capture(op, _lib.ptr(x.v), _lib.ptr(y.v), stream))
        A._graphs["bound"] = key
    _lib.check(L.hpcla_spmv_graph_launch(op, stream))
    return y


def enable_direct_halo(A: HPCSparseMatrix, x: HPCVector) -> None:
    """Switch the halo exchange of (A, x.partition) from grouped ncclSend/ncclRecv to the direct push: every rank maps its
    neighbours' `gathered` buffers (CUDA IPC between processes, plain pointers between rank-threads) and from then on
    copies its ghost runs straight into them over NVLink with the copy engine, announcing them with flags the receiving
    stream waits on (no kernel spins).  Collective: every rank exports a blob, the host communicator all-gathers them
    (like the 128-byte NCCL id of ext/HPCLinearAlgebraCUDAExt.jl:411-443), every rank connects.  NCCL stays the default."""
    L = _lib.lib()
    op = _bound_op(A, get_vector_plan(A, x), x)
    nbytes = ctypes.c_int64()
    _lib.check(L.hpcla_spmv_halo_blob_size(op, ctypes.byref(nbytes)))
    blob = np.zeros(nbytes.value, dtype=np.uint8)
    _lib.check(L.hpcla_spmv_halo_export(op, _lib.ptr(blob)))
    blobs = comm_allgather(A.backend.comm, blob.tobytes())
    allb = np.frombuffer(b"".join(blobs), dtype=np.uint8).copy()
    comm_barrier(A.backend.comm)
    _lib.check(L.hpcla_spmv_halo_connect(op, _lib.ptr(allb)))
    comm_barrier(A.backend.comm)
    A._graphs.clear()


def spmv_timeline(A: HPCSparseMatrix, x: HPCVector) -> Dict[str, float]:
    """Timeline of the most recent multiply (operators created under HPCLA_TIMELINE=1): milliseconds from 'x ready' to
    the end of the halo exchange, of the boundary tiles, of the interior tiles and of the call."""
    op = _bound_op(A, get_vector_plan(A, x), x)
    out = (ctypes.c_double * 4)()
    _lib.check(_lib.lib().hpcla_spmv_timeline(op, out))
    return {"exchange_ms": out[0], "boundary_ms": out[1], "interior_ms": out[2], "end_ms": out[3]}


def mul_staged(y: HPCVector, A: HPCSparseMatrix, x: HPCVector, x_host, y_host) -> HPCVector:
    """copyto!(x.v, x_host); mul!(y, A, x); copyto!(y_host, y.v) as ONE pipelined call (hpcla_spmv_run_staged).

    This is the reference's CUDA path seen from the host — every `A * x` there stages x through the host and copies
    the gathered vector back up (src/vectors.jl:423, 460), and mul! computes on host arrays (src/sparse.jl:2019-2037)
    — with the upload, the multiply and the download overlapped block by block.  x_host / y_host: this rank's local
    slices as (pinned) CPU torch tensors or numpy arrays of A's element type.  y_host is complete after the current
    stream has been synchronised."""
    _check_mul_args(A, x)
    if y.local_size != A.nrows_local:
        raise ValueError(f"DimensionMismatch: y holds {y.local_size} local rows, A has {A.nrows_local}")
    if A.backend.ctx().world == "threads":
        raise _lib.HPCLAError("mul_staged needs an NCCL world or a single rank")
    for name, h, n in (("x_host", x_host, x.local_size), ("y_host", y_host, y.local_size)):
        hn = h.numel() if hasattr(h, "numel") else h.size
        if hn != n:
            raise ValueError(f"DimensionMismatch: {name} holds {hn} elements, the local slice has {n}")
        if hasattr(h, "is_cuda") and h.is_cuda:
            raise ValueError(f"{name} must be a host array")
    op = _bound_op(A, get_vector_plan(A, x), x)
    _lib.check(_lib.lib().hpcla_spmv_run_staged(op, _lib.ptr(x_host), _lib.ptr(x.v), _lib.ptr(y.v), _lib.ptr(y_host), _current_stream(A.backend)))
    return y


def matvec(A: HPCSparseMatrix, x: HPCVector) -> HPCVector:
    """Base.:*(A, x) — src/sparse.jl:2096-2128.  Result partition = A.row_partition (hash cached in the plan)."""
    import torch

    _check_mul_args(A, x)
    plan = get_vector_plan(A, x)
    if plan.result_partition_hash is None:  # :2103-2106
        plan.result_partition_hash = compute_partition_hash(A.row_partition)
        plan.result_partition = A.row_partition.copy()
    y_local = torch.empty(A.nrows_local, dtype=_torch_dtype(A.backend.T), device=A.backend.torch_device())  # :2115
    _run_multiply(A, x, y_local)
    return HPCVector(plan.result_partition_hash, plan.result_partition, y_local, A.backend)


# ---------------------------------------------------------------------------------------------------------------
# transpose
# ---------------------------------------------------------------------------------------------------------------
class Transpose:
    """Lazy transpose(A) (src/sparse.jl:2254-2260)."""

    def __init__(self, parent: HPCSparseMatrix):
        self.parent = parent

    def __matmul__(self, x):
        from .dense import HPCMatrix, spmm

        if isinstance(x, HPCMatrix):  # transpose(A) * B — src/sparse.jl:2420-2424: materialise A^T, then A^T * B
            return spmm(materialize_transpose(self.parent), x)
        return transpose_matvec(self.parent, x)

    __mul__ = __matmul__

    @property
    def T(self) -> HPCSparseMatrix:
        return self.parent


def transpose(A: HPCSparseMatrix) -> Transpose:
    return Transpose(A)


def materialize_transpose(A: HPCSparseMatrix) -> HPCSparseMatrix:
    """HPCSparseMatrix(transpose(A)) — src/sparse.jl:1846-1865: TransposePlan + execute_plan!, cached bidirectionally."""
    if A.cached_transpose is not None:  # :1849-1851
        return A.cached_transpose
    L = _lib.lib()
    b = A.backend
    comm = b.comm
    rank, P = comm_rank(comm), comm_size(comm)
    if b.is_cuda and os.environ.get("HPCLA_TRANSPOSE", "device") != "host" and (P == 1 or b.ctx().world == "nccl"):
        Y = _materialize_transpose_device(A)
        # The structural hash is a cache key (src/sparse.jl:97-121), and the structure of A^T is a function of the structure
        # of A: derive Y's key from A's instead of hashing ~|A| bytes of read-back arrays a second time on the host (the
        # larger part of the first transpose(A)*x at BASELINE config 3).  Same on every rank (A's hash is the all-gathered
        # one); a matrix with the same structure built another way only misses the plan cache, it can never hit wrongly.
        Y.structural_hash = _digest(b"hpcla:transpose-of:", _ensure_hash(A))
        A.cached_transpose = Y
        Y.cached_transpose = A
        return Y
    nz = np.ascontiguousarray(A.nzval_host())  # _ensure_cpu(A.nzval) (:1760)
    tb = ctypes.c_void_p()
    _lib.check(L.hpcla_transpose_begin(rank, P, _lib.dtype_code(b.T), _lib.itype_code(b.Ti), _lib.ptr(A.row_partition), _lib.ptr(A.col_partition),
                                       _lib.ptr(A.rowptr), _lib.ptr(A.colval), _lib.ptr(np.ascontiguousarray(A.col_indices, dtype=np.int64)), _lib.ptr(nz),
                                       ctypes.byref(tb)))
    try:
        counts = np.zeros(P, dtype=np.int64)
        _lib.check(L.hpcla_tb_counts(tb, _lib.ptr(counts)))
        send_pairs, send_vals = {}, {}
        for q in range(P):
            if q != rank and counts[q] > 0:
                pr = np.empty(2 * int(counts[q]), dtype=np.int64)
                vl = np.empty(int(counts[q]), dtype=b.T)
                _lib.check(L.hpcla_tb_message(tb, q, _lib.ptr(pr), _lib.ptr(vl)))
                send_pairs[q], send_vals[q] = pr, vl
        got_pairs = comm_exchange(comm, send_pairs, np.int64, tag=10)  # structure (:1581-1624)
        got_vals = comm_exchange(comm, send_vals, b.T, tag=11)  # values (:1770-1796)
        rc = np.zeros(P, dtype=np.int64)
        pl, vl = [None] * P, [None] * P
        for q, a in got_pairs.items():
            rc[q] = len(a) // 2
            pl[q] = np.ascontiguousarray(a, dtype=np.int64)
            vl[q] = np.ascontiguousarray(got_vals[q], dtype=b.T)
        nrows, nnz, ncc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.hpcla_transpose_finish(tb, _lib.ptr(rc), _lib.ptr_array(pl), _lib.ptr_array(vl), ctypes.byref(nrows), ctypes.byref(nnz), ctypes.byref(ncc)))
        rowptr = np.empty(nrows.value + 1, dtype=b.Ti)
        colval = np.empty(nnz.value, dtype=b.Ti)
        col_indices = np.empty(ncc.value, dtype=np.int64)
        nzval = np.empty(nnz.value, dtype=b.T)
        _lib.check(L.hpcla_tb_result(tb, _lib.ptr(rowptr), _lib.ptr(colval), _lib.ptr(col_indices), _lib.ptr(nzval)))
    finally:
        L.hpcla_tb_destroy(tb)
    # row_partition = A.col_partition, col_partition = A.row_partition (:1560-1561)
    Y = HPCSparseMatrix(None, A.col_partition.copy(), A.row_partition.copy(), col_indices, rowptr, colval, _to_device(nzval, b), int(nrows.value),
                        int(ncc.value), _to_device(rowptr, b), _to_device(colval, b), b)
    A.cached_transpose = Y  # :1858-1859
    Y.cached_transpose = A
    return Y


def _materialize_transpose_device(A: HPCSparseMatrix) -> HPCSparseMatrix:
    """The same result without leaving the GPU (hpcla_transpose_device): keys built on the device,
    exchanged over NCCL, radix-sorted, turned into CSR.  The host copies of rowptr / colval the structure-side code
    expects (plans, hashes) are read back once."""
    import torch

    L = _lib.lib()
    b = A.backend
    h = ctypes.c_void_p()
    ci = np.ascontiguousarray(A.col_indices, dtype=np.int64)
    stream = _current_stream(b)
    _lib.check(L.hpcla_transpose_device(b.ctx().handle, _lib.dtype_code(b.T), _lib.itype_code(b.Ti), _lib.ptr(A.row_partition), _lib.ptr(A.col_partition),
                                        A.nrows_local, A.ncols_compressed, A.nnz_local, _lib.ptr(A.rowptr_target), _lib.ptr(A.colval_target), _lib.ptr(ci),
                                        _lib.ptr(A.nzval), stream, ctypes.byref(h)))
    try:
        nrows, nnz, ncc = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.hpcla_dtb_sizes(h, ctypes.byref(nrows), ctypes.byref(nnz), ctypes.byref(ncc)))
        dev = b.torch_device()
        ti = torch.int32 if np.dtype(b.Ti) == np.int32 else torch.int64
        rowptr_d = torch.empty(nrows.value + 1, dtype=ti, device=dev)
        colval_d = torch.empty(nnz.value, dtype=ti, device=dev)
        nzval_d = torch.empty(nnz.value, dtype=_torch_dtype(b.T), device=dev)
        col_indices = np.empty(ncc.value, dtype=np.int64)
        _lib.check(L.hpcla_dtb_result(h, _lib.ptr(rowptr_d), _lib.ptr(colval_d), _lib.ptr(col_indices), _lib.ptr(nzval_d), stream))
    finally:
        L.hpcla_dtb_destroy(h)
    return HPCSparseMatrix(None, A.col_partition.copy(), A.row_partition.copy(), col_indices, rowptr_d.cpu().numpy(), colval_d.cpu().numpy(), nzval_d,
                           int(nrows.value), int(ncc.value), rowptr_d, colval_d, b)


def transpose_matvec(A: HPCSparseMatrix, x: HPCVector) -> HPCVector:
    """Base.:*(transpose(A), x) — src/sparse.jl:2375-2379.  Result partition = A.col_partition; never conjugates."""
    return matvec(materialize_transpose(A), x)


def vec_transpose_mul(x: HPCVector, A: HPCSparseMatrix) -> HPCVector:
    """transpose(v) * A = transpose(transpose(A) * v) — src/sparse.jl:2136-2142 (returned as the untransposed vector)."""
    return transpose_matvec(A, x)


def vec_adjoint_mul(x: HPCVector, A: HPCSparseMatrix) -> HPCVector:
    """v' * A = transpose(conj(v)) * A — src/vectors.jl:746 routed through src/sparse.jl:2136-2142."""
    return transpose_matvec(A, x.conj())


# ---------------------------------------------------------------------------------------------------------------
# to_backend (src/HPCLinearAlgebra.jl:337-378)
# ---------------------------------------------------------------------------------------------------------------
def to_backend(obj, backend: HPCBackend):
    if isinstance(obj, HPCVector):
        return HPCVector(obj.structural_hash, obj.partition, _to_device(obj.local_values().astype(backend.T, copy=False), backend), backend)
    if isinstance(obj, HPCSparseMatrix):
        A = obj
        return HPCSparseMatrix(A.structural_hash, A.row_partition, A.col_partition, A.col_indices, A.rowptr.astype(backend.Ti), A.colval.astype(backend.Ti),
                               _to_device(A.nzval_host().astype(backend.T, copy=False), backend), A.nrows_local, A.ncols_compressed,
                               _to_device(A.rowptr.astype(backend.Ti), backend), _to_device(A.colval.astype(backend.Ti), backend), backend)
    raise TypeError(f"to_backend: unsupported object {type(obj)}")


# ---------------------------------------------------------------------------------------------------------------
# CG (a user-level composition in the reference; here one library call, no host sync per iteration)
# ---------------------------------------------------------------------------------------------------------------
def cg(A: HPCSparseMatrix, b: HPCVector, iters: int, x: Optional[HPCVector] = None, work=None):
    """Fixed-iteration conjugate gradients, x0 = 0.  Returns (x, rr_history) with rr_history[k] = dot(r, r) after
    iteration k+1.  x / work (3 * local_size elements of device scratch): optional caller-owned buffers — with
    HPCLA_CG_GRAPH=1 the whole loop is one CUDA graph that is re-used as long as b, x, work and iters stay the same."""
    import torch

    _check_mul_args(A, b)
    plan = get_vector_plan(A, b)
    op = _bound_op(A, plan, b)
    if x is None:
        x = b.similar()
    if work is None:
        work = torch.empty(3 * b.local_size, dtype=_torch_dtype(A.backend.T), device=A.backend.torch_device())
    hist = np.zeros(max(iters, 1), dtype=np.float64)
    _lib.check(_lib.lib().hpcla_cg(op, _lib.ptr(b.v), _lib.ptr(x.v), _lib.ptr(work), int(iters), _lib.ptr(hist), _current_stream(A.backend)))
    return x, hist[:iters]
