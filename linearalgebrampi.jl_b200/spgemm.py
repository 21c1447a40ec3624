"""Sparse x sparse, `A * B` for two HPCSparseMatrix — mirror of src/sparse.jl:554-1059.

`MatrixPlan(A, B)` (src/sparse.jl:579-922) gathers the structure of the rows of B that A's columns reference
(`A.col_indices`), once per pair of structures (memoised like the reference's `_plan_cache`, :900-916).  The reference
then redoes the whole sparse product on the host at every call (`plan.AT * A_csc`, :1011).  Here the plan also holds the
memoised SYMBOLIC product (hpcla_spgemm_symbolic: structure of C and, per stored entry of C, the pairs of stored entries
that feed it), so a multiply with new values is: pack the requested values of B (device), one exchange of byte ranges
between devices (hpcla_exchange_bytes, no host staging), one kernel (hpcla_spgemm_numeric).
"""
from __future__ import annotations

import ctypes
from typing import Dict

import numpy as np

from . import _lib
from .backends import HPCBackend, assert_backends_compatible, comm_allgather, comm_barrier, comm_exchange, comm_rank, comm_size
from .sparse import HPCSparseMatrix, _comm_key, _ensure_hash
from .vectors import _current_stream, _torch_dtype

_matrix_plan_cache: Dict[tuple, "MatrixPlan"] = {}
matrix_plan_build_count = 0


def _owner(partition: np.ndarray, g: np.ndarray) -> np.ndarray:
    """Owner of 1-based global indices in a partition of 1-based starts (searchsortedlast - 1, clamped: :598-606)."""
    o = np.searchsorted(partition, g, side="right") - 1
    return np.clip(o, 0, len(partition) - 2)


class MatrixPlan:
    """MatrixPlan{T,Ti,AIV} (src/sparse.jl:554-565) + the memoised symbolic product.

    Index fields (host): `bg_rowptr` / `bg_cols` = structure of B[A.col_indices, :] with GLOBAL columns (the pattern of
    plan.AT); `send_pos[q]` = positions in B.nzval of the rows rank q asked for (send_ranges, flattened);
    `recv_off[q]` / `recv_cnt[q]` = where owner q's values land in the gathered value array (recv_offsets)."""

    def __init__(self, A: HPCSparseMatrix, B: HPCSparseMatrix):
        b = A.backend
        comm = b.comm
        me, P = comm_rank(comm), comm_size(comm)
        rpB = np.ascontiguousarray(B.row_partition, dtype=np.int64)
        need = np.ascontiguousarray(A.col_indices, dtype=np.int64)  # sorted global rows of B (:900-916 passes A.col_indices)
        if len(need) and (need[0] < 1 or need[-1] > rpB[-1] - 1):
            raise ValueError(f"DimensionMismatch: A has columns up to {int(need[-1])}, B has {int(rpB[-1]) - 1} rows")
        own = _owner(rpB, need)
        # step 1 (:608-640): ask every owner for the structure of the rows it holds
        seg = np.searchsorted(own, np.arange(P + 1))  # owners ascend with the sorted row ids: one contiguous run each
        requests = {q: need[seg[q]:seg[q + 1]] for q in range(P) if q != me and seg[q + 1] > seg[q]}
        wanted = comm_exchange(comm, requests, np.int64, tag=1)
        wanted[me] = need[seg[me]:seg[me + 1]]
        # step 2 (:642-760): answer with row lengths and GLOBAL column ids; remember which values each requester gets
        Brp = np.asarray(B.rowptr, dtype=np.int64)
        Bcv = np.asarray(B.colval, dtype=np.int64)
        Bci = np.asarray(B.col_indices, dtype=np.int64)
        replies, self.send_pos = {}, {}
        for q, rows in wanted.items():
            loc = np.asarray(rows, dtype=np.int64) - rpB[me]  # 0-based local rows
            if len(loc) and (loc.min() < 0 or loc.max() >= B.nrows_local):
                raise _lib.HPCLAError(f"MatrixPlan: rank {q} asked rank {me} for a row it does not own")
            beg, end = Brp[loc] - 1, Brp[loc + 1] - 1
            lens = end - beg
            total = int(lens.sum())
            pos = np.arange(total, dtype=np.int64) + np.repeat(beg - (np.cumsum(lens) - lens), lens)  # beg[r] .. end[r]-1 for every row, concatenated
            self.send_pos[q] = pos
            replies[q] = np.concatenate([lens, Bci[Bcv[pos] - 1]]).astype(np.int64)
        got = comm_exchange(comm, {q: a for q, a in replies.items() if q != me}, np.int64, tag=2)
        got[me] = replies.get(me, np.zeros(0, np.int64))
        # step 3 (:762-897): the pattern of B[need, :]
        lens_all, cols_all = [], []
        self.recv_off, self.recv_cnt = np.zeros(P, np.int64), np.zeros(P, np.int64)
        off = 0
        for q in range(P):
            nrows_q = int(seg[q + 1] - seg[q])
            a = got.get(q, np.zeros(0, np.int64))
            lens_q, cols_q = a[:nrows_q], a[nrows_q:]
            if len(lens_q) != nrows_q or int(lens_q.sum()) != len(cols_q):
                raise _lib.HPCLAError(f"MatrixPlan: malformed structure reply from rank {q}")
            lens_all.append(lens_q)
            cols_all.append(cols_q)
            self.recv_off[q], self.recv_cnt[q] = off, len(cols_q)
            off += len(cols_q)
        lens_all = np.concatenate(lens_all) if lens_all else np.zeros(0, np.int64)
        self.bg_rowptr = np.concatenate([[1], 1 + np.cumsum(lens_all)]).astype(np.int64)
        self.bg_cols = np.ascontiguousarray(np.concatenate(cols_all) if cols_all else np.zeros(0, np.int64), dtype=np.int64)
        self.n_gathered_values = int(off)
        # memoised symbolic product
        L = _lib.lib()
        h = ctypes.c_void_p()
        rc = L.hpcla_spgemm_symbolic(_lib.itype_code(b.Ti), A.nrows_local, _lib.ptr(np.ascontiguousarray(A.rowptr)), _lib.ptr(np.ascontiguousarray(A.colval)),
                                     len(need), _lib.ptr(self.bg_rowptr), _lib.ptr(self.bg_cols), ctypes.byref(h))
        # a rank-local failure (out of memory for the term lists) must not leave the other ranks waiting in the value exchange
        # of the first product: every rank learns the worst status before anyone goes on
        msg = (L.hpcla_last_error() or b"").decode() if rc else ""
        statuses = comm_allgather(b.comm, (int(rc), msg))
        bad = [(q, st) for q, st in enumerate(statuses) if st[0] != 0]
        if bad:
            if h.value:
                L.hpcla_spgemm_destroy(h)
            raise _lib.HPCLAError("MatrixPlan: the symbolic product failed on rank(s) " + ", ".join(f"{q} [status {st[0]}: {st[1]}]" for q, st in bad))
        self.handle = h.value
        nnz, ncc, nterms = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.hpcla_spgemm_sizes(self.handle, ctypes.byref(nnz), ctypes.byref(ncc), ctypes.byref(nterms)))
        self.nnz, self.ncc, self.nterms = nnz.value, ncc.value, nterms.value
        self.rowptr = np.empty(A.nrows_local + 1, dtype=b.Ti)
        self.colval = np.empty(self.nnz, dtype=b.Ti)
        self.col_indices = np.empty(self.ncc, dtype=np.int64)
        _lib.check(L.hpcla_spgemm_structure(self.handle, _lib.itype_code(b.Ti), _lib.ptr(self.rowptr), _lib.ptr(self.colval), _lib.ptr(self.col_indices)))
        self._dev = None  # device copy of the send positions, made on first execution
        self._structure_dev = None  # device copies of C's rowptr / colval, shared by every product of this plan

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().hpcla_spgemm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def get_matrix_plan(A: HPCSparseMatrix, B: HPCSparseMatrix) -> MatrixPlan:
    """MatrixPlan(A, B) — src/sparse.jl:900-916: memoised on (hash(A), hash(B), T, Ti, backend)."""
    global matrix_plan_build_count
    key = (_ensure_hash(A), _ensure_hash(B), A.backend.T.str, A.backend.Ti.str, _comm_key(A.backend))
    plan = _matrix_plan_cache.get(key)
    if plan is None:
        plan = _matrix_plan_cache[key] = MatrixPlan(A, B)
        matrix_plan_build_count += 1
    return plan


def _gather_values(plan: MatrixPlan, B: HPCSparseMatrix):
    """execute_plan!(plan, B) — src/sparse.jl:922-983 — on the device: B.nzval[send positions] packed per requester, one
    exchange of byte ranges, the values of B[A.col_indices, :] in gathered order."""
    import torch

    b = B.backend
    comm = b.comm
    me, P = comm_rank(comm), comm_size(comm)
    dev = b.torch_device()
    es = np.dtype(b.T).itemsize
    if plan._dev is None:
        order = sorted(plan.send_pos)
        pos = np.concatenate([plan.send_pos[q] for q in order]) if order else np.zeros(0, np.int64)
        send_off, send_cnt = np.zeros(P, np.int64), np.zeros(P, np.int64)
        off = 0
        for q in order:
            send_off[q], send_cnt[q] = off, len(plan.send_pos[q])
            off += len(plan.send_pos[q])
        plan._dev = dict(pos=torch.from_numpy(pos).to(dev), send_off=send_off, send_cnt=send_cnt)
    d = plan._dev
    packed = B.nzval.index_select(0, d["pos"]) if len(d["pos"]) else torch.empty(0, dtype=_torch_dtype(b.T), device=dev)
    out = torch.empty(plan.n_gathered_values, dtype=_torch_dtype(b.T), device=dev)
    ctx = b.ctx()
    if ctx.world == "threads":
        # single-process harness world: every rank publishes its packed values, the receivers copy their ranges
        pub = comm_allgather(comm, (packed, d["send_off"], d["send_cnt"]))
        for q in range(P):
            n = int(plan.recv_cnt[q])
            if n == 0:
                continue
            src, soff, scnt = pub[q]
            if int(scnt[me]) != n:
                raise _lib.HPCLAError(f"MatrixPlan: rank {q} packs {int(scnt[me])} values for rank {me}, which expects {n}")
            out[int(plan.recv_off[q]) : int(plan.recv_off[q]) + n].copy_(src[int(soff[me]) : int(soff[me]) + n])
        torch.cuda.synchronize(dev)
        comm_barrier(comm)  # nobody drops its packed values before every reader is done
    else:
        so, sb = np.ascontiguousarray(d["send_off"] * es), np.ascontiguousarray(d["send_cnt"] * es)
        ro, rb = np.ascontiguousarray(plan.recv_off * es), np.ascontiguousarray(plan.recv_cnt * es)
        _lib.check(_lib.lib().hpcla_exchange_bytes(ctx.handle, _lib.ptr(packed), _lib.ptr(so), _lib.ptr(sb), _lib.ptr(out), _lib.ptr(ro), _lib.ptr(rb),
                                                   _current_stream(b)))
        out._hpcla_keepalive = packed  # read by the enqueued exchange
    return out


def spgemm(A: HPCSparseMatrix, B: HPCSparseMatrix) -> HPCSparseMatrix:
    """Base.:*(A::HPCSparseMatrix, B::HPCSparseMatrix) — src/sparse.jl:991-1059.  Result: row_partition = A's,
    col_partition = B's, col_indices = unique(sort(columns)), values in the reference's summation order."""
    import torch

    b = A.backend
    assert_backends_compatible(b, B.backend)
    if not b.is_cuda:
        raise _lib.HPCLAError("A*B needs DeviceCUDA operands: this build has no CPU arithmetic (and no CPU fallback)")
    if B.backend.T != b.T or B.backend.Ti != b.Ti:
        raise TypeError("A and B must share element and index types")
    if A.shape[1] != B.shape[0]:
        raise ValueError(f"DimensionMismatch: A has {A.shape[1]} columns, B has {B.shape[0]} rows")
    plan = get_matrix_plan(A, B)
    bg_vals = _gather_values(plan, B)
    dev = b.torch_device()
    nzval = torch.empty(plan.nnz, dtype=_torch_dtype(b.T), device=dev)
    _lib.check(_lib.lib().hpcla_spgemm_numeric(plan.handle, b.ctx().handle, _lib.dtype_code(b.T), _lib.ptr(A.nzval), _lib.ptr(bg_vals), _lib.ptr(nzval),
                                               _current_stream(b)))
    nzval._hpcla_keepalive = bg_vals
    # the structure of C is part of the memoised plan: every product shares the same (read-only) structure arrays, on the
    # host and on the device; only nzval is new
    if plan._structure_dev is None:
        plan._structure_dev = (torch.from_numpy(plan.rowptr).to(dev), torch.from_numpy(plan.colval).to(dev))
    rowptr_t, colval_t = plan._structure_dev
    return HPCSparseMatrix(None, A.row_partition.copy(), B.col_partition.copy(), plan.col_indices, plan.rowptr, plan.colval, nzval,
                           A.nrows_local, plan.ncc, rowptr_t, colval_t, b)
