"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of the C++ CPU restatement (oracle/hpcla_oracle.cpp) of the reference's distributed sparse
mat-vec path, plus an independent pure-numpy twin (`np_*` functions) used to cross-check the C++ restatement.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Parity status: arithmetic pinned by the reference's own test fixtures (tests/golden/); plan arrays and ghost maps are
read by no reference test, so for them parity is unpinned by the reference (see hpcla_oracle.cpp header).

All index arrays are 1-based, exactly as the reference stores them (src/sparse.jl:319-337, src/vectors.jl:229-251).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhpcla_oracle.so")

DTYPES = {"f32": np.dtype(np.float32), "f64": np.dtype(np.float64), "c128": np.dtype(np.complex128)}
DTYPE_CODE = {"f32": 0, "f64": 1, "c128": 2}
ITYPES = {"i32": np.dtype(np.int32), "i64": np.dtype(np.int64)}
ITYPE_CODE = {"i32": 0, "i64": 1}


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "hpcla_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libhpcla_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        L.orc_uniform_partition.argtypes = [i64, i64, vp]
        L.orc_owner.argtypes = [vp, i64, i64]
        L.orc_owner.restype = i64
        L.orc_compress.argtypes = [i64, vp, vp, vp]
        L.orc_compress.restype = i64
        L.orc_plans_build.argtypes = [i64, vp, vp, vp]
        L.orc_plans_build.restype = vp
        L.orc_plans_free.argtypes = [vp]
        L.orc_plans_len.argtypes = [vp, i64, ci, i64]
        L.orc_plans_len.restype = i64
        L.orc_plans_get.argtypes = [vp, i64, ci, i64, vp]
        L.orc_plans_execute.argtypes = [vp, vp, vp, i64]
        for name in ("f32_i32", "f32_i64", "f64_i32", "f64_i64", "c128_i32", "c128_i64"):
            getattr(L, "orc_spmv_" + name).argtypes = [i64, vp, vp, vp, vp, vp]
        L.orc_dot_f64.argtypes = [i64, vp, vp]
        L.orc_dot_f64.restype = ctypes.c_double
        L.orc_dot_f32.argtypes = [i64, vp, vp]
        L.orc_dot_f32.restype = ctypes.c_float
        L.orc_dot_c128.argtypes = [i64, vp, vp, vp]
        L.orc_transpose_build.argtypes = [i64, vp, vp, vp, vp, vp, vp, i64]
        L.orc_transpose_build.restype = vp
        L.orc_transpose_free.argtypes = [vp]
        L.orc_transpose_len.argtypes = [vp, i64, ci]
        L.orc_transpose_len.restype = i64
        L.orc_transpose_get.argtypes = [vp, i64, ci, vp]
        L.orc_bench_spmv.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp, vp, i64, i64, vp]
        L.orc_bench_spmv.restype = ci
        u64 = ctypes.c_uint64
        L.orc_bench_synth.argtypes = [vp, ci, i64, i64, i64, u64, i64, u64, ci, ci, i64, ci, ci, i64, i64, i64, vp, vp, vp]
        L.orc_bench_synth.restype = ci
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _ptr_array(arrs: Sequence[np.ndarray]):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def dtype_name(dt) -> str:
    dt = np.dtype(dt)
    for k, v in DTYPES.items():
        if v == dt:
            return k
    raise ValueError(f"unsupported element type {dt}")


def itype_name(it) -> str:
    it = np.dtype(it)
    for k, v in ITYPES.items():
        if v == it:
            return k
    raise ValueError(f"unsupported index type {it}")


# ------------------------------------------------------------------------------------------------------------------
# containers
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class LocalMatrix:
    """The fields of one rank's HPCSparseMatrix (src/sparse.jl:319-337)."""

    rank: int
    row_partition: np.ndarray
    col_partition: np.ndarray
    col_indices: np.ndarray  # int64, sorted global columns present locally
    rowptr: np.ndarray  # Ti, 1-based, nrows_local+1
    colval: np.ndarray  # Ti, 1-based LOCAL index into col_indices
    nzval: np.ndarray
    nrows_local: int
    ncols_compressed: int


@dataclass
class Plan:
    """The index fields of one rank's VectorPlan (src/vectors.jl:229-251)."""

    send_rank_ids: np.ndarray
    send_indices: List[np.ndarray]
    recv_rank_ids: np.ndarray
    recv_perm: List[np.ndarray]
    local_src_indices: np.ndarray
    local_dst_indices: np.ndarray
    n_gathered: int


# ------------------------------------------------------------------------------------------------------------------
# C++-backed restatement
# ------------------------------------------------------------------------------------------------------------------
def uniform_partition(n: int, nranks: int) -> np.ndarray:
    out = np.empty(nranks + 1, dtype=np.int64)
    lib().orc_uniform_partition(int(n), int(nranks), _p(out))
    return out


def owner(partition: np.ndarray, g: int) -> int:
    partition = np.ascontiguousarray(partition, dtype=np.int64)
    return int(lib().orc_owner(_p(partition), len(partition) - 1, int(g)))


def compress(global_cols: np.ndarray):
    """(col_indices, colval) per src/sparse.jl:501 and :137-144; inputs/outputs 1-based int64."""
    g = np.ascontiguousarray(global_cols, dtype=np.int64)
    ci = np.empty(len(g), dtype=np.int64)
    cv = np.empty(len(g), dtype=np.int64)
    ncc = lib().orc_compress(len(g), _p(g), _p(ci), _p(cv))
    return ci[:ncc].copy(), cv


def local_matrix(rank, rowptr1, global_cols1, nzval, row_partition, col_partition, itype="i64") -> LocalMatrix:
    """HPCSparseMatrix_local (src/sparse.jl:454-525) given this rank's rows with GLOBAL 1-based columns."""
    it = ITYPES[itype]
    ci, cv = compress(global_cols1)
    return LocalMatrix(
        rank=rank,
        row_partition=np.asarray(row_partition, dtype=np.int64).copy(),
        col_partition=np.asarray(col_partition, dtype=np.int64).copy(),
        col_indices=ci,
        rowptr=np.ascontiguousarray(rowptr1, dtype=it),
        colval=cv.astype(it),
        nzval=np.ascontiguousarray(nzval),
        nrows_local=len(rowptr1) - 1,
        ncols_compressed=len(ci),
    )


def distribute(A, nranks, row_partition=None, col_partition=None, itype="i64", dtype=None) -> List[LocalMatrix]:
    """HPCSparseMatrix{T}(A_global, backend; row_partition, col_partition) on every rank (src/sparse.jl:398-413).

    `A` is a scipy.sparse matrix; it is canonicalised like Julia's `sparse(I,J,V)` (duplicates summed, explicit
    zeros kept, columns ascending within a row)."""
    import scipy.sparse as sp

    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    m, n = A.shape
    rp = uniform_partition(m, nranks) if row_partition is None else np.asarray(row_partition, dtype=np.int64)
    cp = uniform_partition(n, nranks) if col_partition is None else np.asarray(col_partition, dtype=np.int64)
    vals = A.data if dtype is None else A.data.astype(DTYPES[dtype] if isinstance(dtype, str) else dtype)
    out = []
    for r in range(nranks):
        r0, r1 = int(rp[r]) - 1, int(rp[r + 1]) - 1
        lo, hi = int(A.indptr[r0]), int(A.indptr[r1])
        rowptr1 = (A.indptr[r0 : r1 + 1].astype(np.int64) - lo) + 1
        gcols1 = A.indices[lo:hi].astype(np.int64) + 1
        out.append(local_matrix(r, rowptr1, gcols1, vals[lo:hi].copy(), rp, cp, itype))
    return out


def vector_plans(locals_: Sequence[LocalMatrix], x_partition) -> List[Plan]:
    """VectorPlan(A, x) for every rank (src/sparse.jl:1875-1984)."""
    L = lib()
    P = len(locals_)
    xp = np.ascontiguousarray(x_partition, dtype=np.int64)
    cis = [np.ascontiguousarray(m.col_indices, dtype=np.int64) for m in locals_]
    ncc = np.array([len(c) for c in cis], dtype=np.int64)
    W = L.orc_plans_build(P, _ptr_array(cis), _p(ncc), _p(xp))
    try:
        plans = []
        for r in range(P):

            def get(field, slot=0):
                n = L.orc_plans_len(W, r, field, slot)
                a = np.empty(n, dtype=np.int64)
                L.orc_plans_get(W, r, field, slot, _p(a))
                return a

            srk, rrk = get(0), get(1)
            plans.append(
                Plan(
                    send_rank_ids=srk,
                    send_indices=[get(4, i) for i in range(len(srk))],
                    recv_rank_ids=rrk,
                    recv_perm=[get(5, i) for i in range(len(rrk))],
                    local_src_indices=get(2),
                    local_dst_indices=get(3),
                    n_gathered=int(ncc[r]),
                )
            )
    finally:
        L.orc_plans_free(W)
    return plans


class PlanWorld:
    """Keeps the C++ plan object alive for execute / bench."""

    def __init__(self, locals_: Sequence[LocalMatrix], x_partition):
        L = lib()
        self.P = len(locals_)
        xp = np.ascontiguousarray(x_partition, dtype=np.int64)
        self._cis = [np.ascontiguousarray(m.col_indices, dtype=np.int64) for m in locals_]
        self.ncc = np.array([len(c) for c in self._cis], dtype=np.int64)
        self.handle = L.orc_plans_build(self.P, _ptr_array(self._cis), _p(self.ncc), _p(xp))

    def execute(self, x_locals: Sequence[np.ndarray]) -> List[np.ndarray]:
        """execute_plan! on every rank (src/vectors.jl:394-463) -> gathered per rank."""
        dt = x_locals[0].dtype
        xs = [np.ascontiguousarray(x, dtype=dt) for x in x_locals]
        gs = [np.zeros(int(n), dtype=dt) for n in self.ncc]
        lib().orc_plans_execute(self.handle, _ptr_array(xs), _ptr_array(gs), dt.itemsize)
        return gs

    def close(self):
        if self.handle:
            lib().orc_plans_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spmv_local(m: LocalMatrix, gathered: np.ndarray) -> np.ndarray:
    """_spmv_kernel! (src/sparse.jl:2055-2066): row-serial, left-to-right, no FMA."""
    name = f"orc_spmv_{dtype_name(m.nzval.dtype)}_{itype_name(m.rowptr.dtype)}"
    y = np.empty(m.nrows_local, dtype=m.nzval.dtype)
    g = np.ascontiguousarray(gathered, dtype=m.nzval.dtype)
    getattr(lib(), name)(m.nrows_local, _p(m.rowptr), _p(m.colval), _p(m.nzval), _p(g), _p(y))
    return y


def spmv_csr(rowptr, colval, nzval, x) -> np.ndarray:
    """The same row loop on bare arrays: y[r] = sum_j nzval[j] * x[colval[j]] (1-based rowptr / colval), left to right."""
    rowptr = np.ascontiguousarray(rowptr)
    colval = np.ascontiguousarray(colval, dtype=rowptr.dtype)
    nzval = np.ascontiguousarray(nzval)
    name = f"orc_spmv_{dtype_name(nzval.dtype)}_{itype_name(rowptr.dtype)}"
    n = len(rowptr) - 1
    y = np.empty(n, dtype=nzval.dtype)
    xg = np.ascontiguousarray(x, dtype=nzval.dtype)
    getattr(lib(), name)(n, _p(rowptr), _p(colval), _p(nzval), _p(xg), _p(y))
    return y


def split_vector(x_global: np.ndarray, partition) -> List[np.ndarray]:
    """HPCVector(v_global, backend; partition) local slices (src/vectors.jl:116-129)."""
    p = np.asarray(partition, dtype=np.int64)
    return [np.ascontiguousarray(x_global[int(p[r]) - 1 : int(p[r + 1]) - 1]) for r in range(len(p) - 1)]


def matvec(locals_: Sequence[LocalMatrix], x_global: np.ndarray, x_partition=None) -> np.ndarray:
    """mul!(y, A, x) / A*x on all ranks, result gathered in row order (src/sparse.jl:2019-2037, 2096-2128)."""
    P = len(locals_)
    xp = locals_[0].col_partition if x_partition is None else np.asarray(x_partition, dtype=np.int64)
    xs = split_vector(np.asarray(x_global, dtype=locals_[0].nzval.dtype), xp)
    W = PlanWorld(locals_, xp)
    try:
        gs = W.execute(xs)
    finally:
        W.close()
    return np.concatenate([spmv_local(locals_[r], gs[r]) for r in range(P)])


def repartition_plan(rank: int, old_partition, new_partition) -> dict:
    """VectorRepartitionPlan(x, p) — src/vectors.jl:519-616, restated loop for loop (1-based values)."""
    xp = np.asarray(old_partition, dtype=np.int64)
    p = np.asarray(new_partition, dtype=np.int64)
    nranks = len(p) - 1
    src_start, src_end = int(xp[rank]), int(xp[rank + 1]) - 1
    dst_start, dst_end = int(p[rank]), int(p[rank + 1]) - 1
    send_ranges_map = {}
    for r in range(nranks):  # :536-552
        r_start, r_end = int(p[r]), int(p[r + 1]) - 1
        if r_end < r_start:
            continue
        o_start, o_end = max(src_start, r_start), min(src_end, r_end)
        if o_start <= o_end:
            send_ranges_map[r] = (o_start - src_start + 1, o_end - src_start + 1)
    # :555-557 — the Alltoall of counts: rank r's count for me is the overlap of ITS source range with MY target range
    recv_counts_raw = []
    for r in range(nranks):
        s_start, s_end = int(xp[r]), int(xp[r + 1]) - 1
        if dst_end < dst_start:
            recv_counts_raw.append(0)
            continue
        o_start, o_end = max(s_start, dst_start), min(s_end, dst_end)
        recv_counts_raw.append(max(0, o_end - o_start + 1))
    local_src, local_dst_offset = (1, 0), 0  # 1:0
    if rank in send_ranges_map:  # :575-582
        local_src = send_ranges_map[rank]
        local_dst_offset = (src_start + local_src[0] - 1) - dst_start + 1
    send_rank_ids = [r for r in range(nranks) if r in send_ranges_map and r != rank]
    send_ranges = [send_ranges_map[r] for r in send_rank_ids]
    recv_rank_ids, recv_counts, recv_offsets = [], [], []
    for r in range(nranks):  # :594-608
        if recv_counts_raw[r] > 0 and r != rank:
            recv_rank_ids.append(r)
            recv_counts.append(recv_counts_raw[r])
            recv_offsets.append(max(int(xp[r]), dst_start) - dst_start + 1)
    return dict(send_rank_ids=send_rank_ids, send_ranges=send_ranges, recv_rank_ids=recv_rank_ids, recv_counts=recv_counts,
                recv_offsets=recv_offsets, local_src_range=local_src, local_dst_offset=local_dst_offset,
                result_local_size=max(0, dst_end - dst_start + 1))


def repartition(xs: Sequence[np.ndarray], old_partition, new_partition) -> List[np.ndarray]:
    """execute_plan!(plan::VectorRepartitionPlan, x) on all ranks — src/vectors.jl:624-676 (messages = array moves)."""
    P = len(xs)
    plans = [repartition_plan(r, old_partition, new_partition) for r in range(P)]
    out = [np.empty(pl["result_local_size"], dtype=xs[0].dtype) for pl in plans]
    mail = {}
    for r, pl in enumerate(plans):
        a, b = pl["local_src_range"]
        if b >= a:
            out[r][pl["local_dst_offset"] - 1 : pl["local_dst_offset"] - 1 + (b - a + 1)] = xs[r][a - 1 : b]
        for dest, (a, b) in zip(pl["send_rank_ids"], pl["send_ranges"]):
            mail[(r, dest)] = xs[r][a - 1 : b].copy()
    for r, pl in enumerate(plans):
        for src, cnt, off in zip(pl["recv_rank_ids"], pl["recv_counts"], pl["recv_offsets"]):
            buf = mail[(src, r)]
            assert len(buf) == cnt
            out[r][off - 1 : off - 1 + cnt] = buf
    return out


def gather_rows(locsB: Sequence[LocalMatrix], row_indices: np.ndarray):
    """What MatrixPlan(row_indices, B) + execute_plan! leave in plan.AT (src/sparse.jl:579-983): the rows
    B[row_indices, :] (sorted global 1-based row ids) as (rowptr 1-based int64, GLOBAL 1-based columns, values), each
    row fetched from its owner in B.row_partition."""
    rpB = np.asarray(locsB[0].row_partition, dtype=np.int64)
    rowptr, cols, vals = [1], [], []
    for g in np.asarray(row_indices, dtype=np.int64):
        o = int(np.searchsorted(rpB, g, side="right")) - 1
        o = min(max(o, 0), len(locsB) - 1)
        m = locsB[o]
        i = int(g - rpB[o])
        b, e = int(m.rowptr[i]) - 1, int(m.rowptr[i + 1]) - 1
        cols.append(np.asarray(m.col_indices, dtype=np.int64)[np.asarray(m.colval[b:e], dtype=np.int64) - 1])
        vals.append(np.asarray(m.nzval[b:e]))
        rowptr.append(rowptr[-1] + (e - b))
    dt = locsB[0].nzval.dtype
    return (np.asarray(rowptr, dtype=np.int64), np.concatenate(cols) if cols else np.zeros(0, np.int64),
            np.concatenate(vals) if vals else np.zeros(0, dt))


def spgemm(locsA: Sequence[LocalMatrix], locsB: Sequence[LocalMatrix], itype="i64") -> List[LocalMatrix]:
    """Base.:*(A::HPCSparseMatrix, B::HPCSparseMatrix) on all ranks (src/sparse.jl:991-1059): gather B[A.col_indices, :],
    then CT = plan.AT * A_csc — SparseArrays' spmatmul: for every local row i of A (a column of A_csc), for its stored
    entries k ascending, for the stored entries (c, b) of gathered row k: CT[c, i] += b * a.  Output columns ascending,
    entries that cancel are kept.  Result block: row_partition = A.row_partition, col_partition = B.col_partition,
    col_indices = unique(sort(columns)), compressed colval (:1018-1040)."""
    out = []
    for A in locsA:
        bg_rowptr, bg_cols, bg_vals = gather_rows(locsB, A.col_indices)
        rowptr, gcols, vals = [1], [], []
        for i in range(A.nrows_local):
            acc, order = {}, []
            for j in range(int(A.rowptr[i]) - 1, int(A.rowptr[i + 1]) - 1):
                g = int(A.colval[j]) - 1
                a = A.nzval[j]
                for q in range(int(bg_rowptr[g]) - 1, int(bg_rowptr[g + 1]) - 1):
                    c = int(bg_cols[q])
                    if c in acc:
                        acc[c] = acc[c] + bg_vals[q] * a
                    else:
                        acc[c] = bg_vals[q] * a
                        order.append(c)
            for c in sorted(order):
                gcols.append(c)
                vals.append(acc[c])
            rowptr.append(rowptr[-1] + len(order))
        dt = A.nzval.dtype
        out.append(local_matrix(A.rank, np.asarray(rowptr, dtype=np.int64), np.asarray(gcols, dtype=np.int64), np.asarray(vals, dtype=dt), A.row_partition,
                                locsB[0].col_partition, itype))
    return out


def to_global(locals_: Sequence[LocalMatrix], shape):
    """The distributed matrix back as one scipy CSR (explicit zeros kept)."""
    import scipy.sparse as sp

    rows, cols, vals = [], [], []
    for m in locals_:
        r0 = int(m.row_partition[m.rank]) - 1
        for i in range(m.nrows_local):
            b, e = int(m.rowptr[i]) - 1, int(m.rowptr[i + 1]) - 1
            rows += [r0 + i] * (e - b)
            cols += (np.asarray(m.col_indices, dtype=np.int64)[np.asarray(m.colval[b:e], dtype=np.int64) - 1] - 1).tolist()
            vals += list(m.nzval[b:e])
    return sp.csr_matrix((np.asarray(vals), (np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64))), shape=shape)


def matmat(locals_: Sequence[LocalMatrix], B_global: np.ndarray, b_row_partition=None) -> np.ndarray:
    """A * B for a dense B, the reference's way (src/sparse.jl:2391-2413): for every column k, `A * B[:, k]` with the
    column's partition = B's row partition (src/indexing.jl:385-393), results concatenated column by column."""
    B_global = np.asarray(B_global)
    cols = [matvec(locals_, B_global[:, k], b_row_partition) for k in range(B_global.shape[1])]
    nrows = int(locals_[0].row_partition[-1]) - 1
    return np.stack(cols, axis=1) if cols else np.zeros((nrows, 0), dtype=locals_[0].nzval.dtype)


def transpose(locals_: Sequence[LocalMatrix]) -> List[LocalMatrix]:
    """HPCSparseMatrix(transpose(A)): TransposePlan + execute_plan! on all ranks (src/sparse.jl:1551-1829)."""
    L = lib()
    P = len(locals_)
    rp = np.ascontiguousarray(locals_[0].row_partition, dtype=np.int64)
    cp = np.ascontiguousarray(locals_[0].col_partition, dtype=np.int64)
    rowptrs = [np.ascontiguousarray(m.rowptr, dtype=np.int64) for m in locals_]
    colvals = [np.ascontiguousarray(m.colval, dtype=np.int64) for m in locals_]
    cis = [np.ascontiguousarray(m.col_indices, dtype=np.int64) for m in locals_]
    nz = [np.ascontiguousarray(m.nzval) for m in locals_]
    dt = nz[0].dtype
    it = locals_[0].rowptr.dtype
    W = L.orc_transpose_build(P, _p(rp), _p(cp), _ptr_array(rowptrs), _ptr_array(colvals), _ptr_array(cis), _ptr_array(nz), dt.itemsize)
    out = []
    try:
        for r in range(P):

            def get(field, dtype=np.int64):
                n = L.orc_transpose_len(W, r, field)
                a = np.empty(n // np.dtype(dtype).itemsize if field == 4 else n, dtype=dtype)
                L.orc_transpose_get(W, r, field, _p(a))
                return a

            rowptr, colval, ci = get(0), get(1), get(2)
            vals = get(4, dt)
            out.append(
                LocalMatrix(
                    rank=r,
                    row_partition=cp.copy(),
                    col_partition=rp.copy(),
                    col_indices=ci,
                    rowptr=rowptr.astype(it),
                    colval=colval.astype(it),
                    nzval=vals,
                    nrows_local=len(rowptr) - 1,
                    ncols_compressed=len(ci),
                )
            )
    finally:
        L.orc_transpose_free(W)
    return out


def dot(x_locals: Sequence[np.ndarray], y_locals: Sequence[np.ndarray]):
    """dot(x, y) = Allreduce(+) of local dots; conjugates x for complex (src/vectors.jl:798-812)."""
    L = lib()
    dt = x_locals[0].dtype
    tot = 0
    for x, y in zip(x_locals, y_locals):
        x = np.ascontiguousarray(x)
        y = np.ascontiguousarray(y, dtype=dt)
        if dt == np.float64:
            tot = tot + L.orc_dot_f64(len(x), _p(x), _p(y))
        elif dt == np.float32:
            tot = np.float32(tot + np.float32(L.orc_dot_f32(len(x), _p(x), _p(y))))
        else:
            o = np.zeros(2)
            L.orc_dot_c128(len(x), _p(x), _p(y), _p(o))
            tot = tot + complex(o[0], o[1])
    return tot


def norm2(x_locals: Sequence[np.ndarray]) -> float:
    """norm(v) = sqrt(Allreduce(+) of local_norm^2) (src/vectors.jl:758-766)."""
    d = dot(x_locals, x_locals)
    return float(np.sqrt(np.real(d)))


def cg(locals_: Sequence[LocalMatrix], b_global: np.ndarray, iters: int, x_partition=None):
    """Textbook CG composed from A*x, dot and axpy exactly as SURVEY §3.5 (the reference ships no CG).
    Returns (x, list of rr = dot(r,r) after every iteration). Fixed iteration count, x0 = 0."""
    b = np.asarray(b_global)
    x = np.zeros_like(b)
    r = b.copy()
    p = r.copy()
    part = locals_[0].row_partition if x_partition is None else x_partition
    rr = dot(split_vector(r, part), split_vector(r, part))
    hist = []
    for _ in range(iters):
        q = matvec(locals_, p, x_partition)
        alpha = rr / dot(split_vector(p, part), split_vector(q, part))
        x = x + alpha * p
        r = r - alpha * q
        rr_new = dot(split_vector(r, part), split_vector(r, part))
        beta = rr_new / rr
        p = r + beta * p
        rr = rr_new
        hist.append(rr)
    return x, hist


def bench_spmv(locals_: Sequence[LocalMatrix], x_locals: Sequence[np.ndarray], x_partition, warmup=1, reps=5):
    """Threaded CPU baseline (one worker thread per rank): returns (times[reps], y_locals)."""
    W = PlanWorld(locals_, x_partition)
    P = W.P
    dt = locals_[0].nzval.dtype
    nrows = np.array([m.nrows_local for m in locals_], dtype=np.int64)
    xs = [np.ascontiguousarray(x, dtype=dt) for x in x_locals]
    ys = [np.zeros(m.nrows_local, dtype=dt) for m in locals_]
    times = np.zeros(reps, dtype=np.float64)
    rc = lib().orc_bench_spmv(
        W.handle,
        DTYPE_CODE[dtype_name(dt)],
        ITYPE_CODE[itype_name(locals_[0].rowptr.dtype)],
        _p(nrows),
        _ptr_array([m.rowptr for m in locals_]),
        _ptr_array([m.colval for m in locals_]),
        _ptr_array([m.nzval for m in locals_]),
        _ptr_array(xs),
        _ptr_array(ys),
        int(warmup),
        int(reps),
        _p(times),
    )
    W.close()
    if rc != 0:
        raise RuntimeError("orc_bench_spmv: unsupported type combination")
    return times, ys


def bench_synth(kind: int, grid, dtype, itype, workers: int, warmup=1, reps=5, op="mul", inner=1, pin=True,
                seed=0xC4, max_len=1_000_000, x_seed=0x5EED):
    """CPU reference arm on a synthetic workload (orc_bench_synth): every worker thread (= one reference MPI rank)
    generates and first-touches its own row block.  kind 0..2: stencil on grid (nx, ny, nz); kind 3: power-law rows,
    n = grid[0].  op "mul": `inner` multiplies per repetition; op "cg": one CG iteration per repetition.
    Returns (times[reps], nnz, sum |y|^2 of the last multiply)."""
    import hpcla_synth  # the generators live in their own library (include/hpcla_synth.h)

    S = hpcla_synth.lib()
    fns = (ctypes.c_void_p * 6)(*[ctypes.cast(getattr(S, n), ctypes.c_void_p).value for n in (
        "hpcla_synth_stencil_nnz", "hpcla_synth_stencil_fill", "hpcla_synth_powerlaw_nnz", "hpcla_synth_powerlaw_fill", "hpcla_synth_vector",
        "hpcla_synth_set_threads")])
    nx, ny, nz = (int(g) for g in grid)
    times = np.zeros(reps, dtype=np.float64)
    nnz = ctypes.c_int64(0)
    yn = ctypes.c_double(0.0)
    rc = lib().orc_bench_synth(fns, int(kind), nx, ny, nz, seed, max_len, x_seed, DTYPE_CODE[dtype_name(dtype)], ITYPE_CODE[itype_name(itype)],
                               int(workers), int(bool(pin)), {"mul": 0, "cg": 1}[op], int(inner), int(warmup), int(reps), _p(times),
                               ctypes.byref(nnz), ctypes.byref(yn))
    if rc != 0:
        raise RuntimeError(f"orc_bench_synth failed (status {rc})")
    return times, int(nnz.value), float(yn.value)


# ------------------------------------------------------------------------------------------------------------------
# Independent numpy twin (second opinion on the C++ restatement; small/medium cases)
# ------------------------------------------------------------------------------------------------------------------
def np_uniform_partition(n, nranks):
    """src/HPCLinearAlgebra.jl:279-289"""
    per, rem = divmod(int(n), int(nranks))
    sizes = np.array([per + (1 if r <= rem else 0) for r in range(1, nranks + 1)], dtype=np.int64)
    return np.concatenate([[1], 1 + np.cumsum(sizes)]).astype(np.int64)


def np_owner(partition, g):
    """searchsortedlast(partition, g) - 1, clamped (src/sparse.jl:1890-1894)"""
    partition = np.asarray(partition)
    o = np.searchsorted(partition, g, side="right") - 1
    return np.minimum(o, len(partition) - 2)


def np_compress(global_cols):
    """src/sparse.jl:501, :137-144"""
    ci = np.unique(np.asarray(global_cols, dtype=np.int64))
    return ci, (np.searchsorted(ci, global_cols, side="left") + 1).astype(np.int64)


def np_vector_plans(col_indices_per_rank, x_partition) -> List[Plan]:
    """src/sparse.jl:1875-1984, written as whole-array numpy instead of per-element pushes."""
    P = len(col_indices_per_rank)
    xp = np.asarray(x_partition, dtype=np.int64)
    owners = [np_owner(xp, np.asarray(ci, dtype=np.int64)) for ci in col_indices_per_rank]
    plans = []
    for r in range(P):
        ci = np.asarray(col_indices_per_rank[r], dtype=np.int64)
        dst = np.arange(1, len(ci) + 1, dtype=np.int64)
        recv_ids = [o for o in range(P) if o != r and np.any(owners[r] == o)]
        send_ids = [q for q in range(P) if q != r and np.any(owners[q] == r)]
        mine = owners[r] == r
        plans.append(
            Plan(
                send_rank_ids=np.array(send_ids, dtype=np.int64),
                send_indices=[
                    np.asarray(col_indices_per_rank[q], dtype=np.int64)[owners[q] == r] - xp[r] + 1 for q in send_ids
                ],
                recv_rank_ids=np.array(recv_ids, dtype=np.int64),
                recv_perm=[dst[owners[r] == o] for o in recv_ids],
                local_src_indices=ci[mine] - xp[r] + 1,
                local_dst_indices=dst[mine],
                n_gathered=len(ci),
            )
        )
    return plans


def np_execute(plans: Sequence[Plan], x_locals: Sequence[np.ndarray]) -> List[np.ndarray]:
    """src/vectors.jl:394-463"""
    P = len(plans)
    gathered = [np.zeros(p.n_gathered, dtype=x_locals[0].dtype) for p in plans]
    mailbox = {}
    for r in range(P):
        p = plans[r]
        gathered[r][p.local_dst_indices - 1] = x_locals[r][p.local_src_indices - 1]
        for i, q in enumerate(p.send_rank_ids):
            mailbox[(r, int(q))] = x_locals[r][p.send_indices[i] - 1]
    for r in range(P):
        p = plans[r]
        for i, q in enumerate(p.recv_rank_ids):
            gathered[r][p.recv_perm[i] - 1] = mailbox[(int(q), r)]
    return gathered


def np_spmv_local(rowptr, colval, nzval, gathered):
    """src/sparse.jl:2055-2066 as a python loop (exact left-to-right order; tiny cases only)."""
    n = len(rowptr) - 1
    y = np.zeros(n, dtype=nzval.dtype)
    for row in range(n):
        acc = nzval.dtype.type(0)
        for j in range(int(rowptr[row]), int(rowptr[row + 1])):
            acc = acc + nzval[j - 1] * gathered[int(colval[j - 1]) - 1]
        y[row] = acc
    return y


def np_transpose(locals_: Sequence[LocalMatrix]) -> List[LocalMatrix]:
    """Result of src/sparse.jl:1551-1829 derived independently: assemble the global matrix, transpose with scipy,
    re-distribute with row_partition = A.col_partition, col_partition = A.row_partition."""
    import scipy.sparse as sp

    rp, cp = locals_[0].row_partition, locals_[0].col_partition
    m, n = int(rp[-1]) - 1, int(cp[-1]) - 1
    rows, cols, vals = [], [], []
    for L_ in locals_:
        r0 = int(rp[L_.rank])
        counts = np.diff(L_.rowptr.astype(np.int64))
        rows.append(np.repeat(np.arange(L_.nrows_local, dtype=np.int64) + (r0 - 1), counts))
        cols.append(L_.col_indices[L_.colval.astype(np.int64) - 1] - 1)
        vals.append(L_.nzval)
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    # coo->csr would DROP nothing and keeps explicit zeros; duplicates cannot exist here
    At = sp.csr_matrix(A.T)
    At.sort_indices()
    return distribute(At, len(locals_), row_partition=cp, col_partition=rp, itype=itype_name(locals_[0].rowptr.dtype))
