"""HPCVector and its ghost/partition maps — host-side mirror of src/vectors.jl for the SpMV hot path.

`HPCVector{T,B}` (src/vectors.jl:21-30): structural_hash (Blake3 of the partition), partition (Int64, 1-based,
length nranks+1), v (the local slice: a torch CUDA tensor on DeviceCUDA backends, a numpy array on DeviceCPU ones),
backend.  Reductions `dot` / `norm` (src/vectors.jl:758-812) and the axpy-type updates needed by iterative solvers
run in libhpcla_b200.so; everything else of vectors.jl (broadcast machinery, repartition) is out of scope.
"""
from __future__ import annotations

import ctypes
import hashlib
from typing import Optional

import numpy as np

from . import _lib
from .backends import (
    HPCBackend,
    assert_backends_compatible,
    comm_allgather,
    comm_allreduce,
    comm_rank,
    comm_size,
)

try:  # the reference's cache keys are Blake3 digests (Blake3Hash.jl); same primitive when available
    from blake3 import blake3 as _blake3

    def _digest(*chunks: bytes) -> bytes:
        h = _blake3()
        for c in chunks:
            h.update(c)
        return h.digest()

except Exception:  # pragma: no cover - hash values are cache keys, not results

    def _digest(*chunks: bytes) -> bytes:
        h = hashlib.blake2b(digest_size=32)
        for c in chunks:
            h.update(c)
        return h.digest()


def uniform_partition(n: int, nranks: int) -> np.ndarray:
    """src/HPCLinearAlgebra.jl:279-289 (computed by the library: hpcla_uniform_partition)."""
    out = np.empty(nranks + 1, dtype=np.int64)
    _lib.check(_lib.lib().hpcla_uniform_partition(int(n), int(nranks), _lib.ptr(out)))
    return out


def compute_partition_hash(partition: np.ndarray) -> bytes:
    """src/HPCLinearAlgebra.jl:255-259: Blake3 of the partition bytes."""
    return _digest(np.ascontiguousarray(partition, dtype=np.int64).tobytes())


def _torch_dtype(T):
    import torch

    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128}[np.dtype(T)]


def _to_device(arr: np.ndarray, backend: HPCBackend):
    """_convert_array / _to_target_device (ext/HPCLinearAlgebraCUDAExt.jl:129-134, 162): host array -> device storage."""
    if not backend.is_cuda:
        return np.ascontiguousarray(arr)
    import torch

    t = torch.from_numpy(np.ascontiguousarray(arr))
    return t.to(backend.torch_device())


def _current_stream(backend: HPCBackend) -> int:
    import torch

    return torch.cuda.current_stream(backend.torch_device()).cuda_stream


class HPCVector:
    """HPCVector{T,B} (src/vectors.jl:21-30)."""

    __slots__ = ("structural_hash", "partition", "v", "backend")

    def __init__(self, structural_hash: bytes, partition: np.ndarray, v, backend: HPCBackend):
        self.structural_hash = structural_hash
        self.partition = partition
        self.v = v
        self.backend = backend

    # -- constructors ------------------------------------------------------------------------------------------
    @staticmethod
    def from_global(v_global, backend: HPCBackend, partition: Optional[np.ndarray] = None) -> "HPCVector":
        """HPCVector(v_global, backend; partition=uniform_partition(...)) — src/vectors.jl:116-129."""
        v_global = np.asarray(v_global)
        P, r = comm_size(backend.comm), comm_rank(backend.comm)
        part = uniform_partition(len(v_global), P) if partition is None else np.ascontiguousarray(partition, dtype=np.int64)
        loc = v_global[int(part[r]) - 1 : int(part[r + 1]) - 1].astype(backend.T, copy=True)
        return HPCVector(compute_partition_hash(part), part, _to_device(loc, backend), backend)

    @staticmethod
    def from_local(v_local, backend: HPCBackend) -> "HPCVector":
        """HPCVector_local(v_local, backend) — src/vectors.jl:76-95: partition from an Allgather of local sizes."""
        n = int(v_local.shape[0])
        sizes = comm_allgather(backend.comm, n)
        part = np.concatenate([[1], 1 + np.cumsum(np.asarray(sizes, dtype=np.int64))]).astype(np.int64)
        if isinstance(v_local, np.ndarray):
            v = _to_device(v_local.astype(backend.T, copy=False), backend)
        else:
            v = v_local  # already device storage
        return HPCVector(compute_partition_hash(part), part, v, backend)

    @staticmethod
    def zeros(backend: HPCBackend, n: int, partition: Optional[np.ndarray] = None) -> "HPCVector":
        """zeros(T, HPCVector, backend, n) — src/HPCLinearAlgebra.jl:1351-1363."""
        P, r = comm_size(backend.comm), comm_rank(backend.comm)
        part = uniform_partition(n, P) if partition is None else np.ascontiguousarray(partition, dtype=np.int64)
        nloc = int(part[r + 1] - part[r])
        if backend.is_cuda:
            import torch

            v = torch.zeros(nloc, dtype=_torch_dtype(backend.T), device=backend.torch_device())
        else:
            v = np.zeros(nloc, dtype=backend.T)
        return HPCVector(compute_partition_hash(part), part, v, backend)

    # -- basics --------------------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return int(self.partition[-1]) - 1

    @property
    def local_size(self) -> int:
        return int(self.v.shape[0])

    def local_values(self) -> np.ndarray:
        """Host copy of the local slice (test/test_utils.jl:235-243 local_values)."""
        return self.v.copy() if isinstance(self.v, np.ndarray) else self.v.detach().cpu().numpy()

    def to_global(self) -> np.ndarray:
        """Vector(v): gather the whole vector on every rank (src/HPCLinearAlgebra.jl:817-840)."""
        parts = comm_allgather(self.backend.comm, self.local_values())
        return np.concatenate(parts) if parts else np.zeros(0, dtype=self.backend.T)

    def similar(self) -> "HPCVector":
        if isinstance(self.v, np.ndarray):
            v = np.empty_like(self.v)
        else:
            import torch

            v = torch.empty_like(self.v)
        return HPCVector(self.structural_hash, self.partition, v, self.backend)

    def copy(self) -> "HPCVector":
        v = self.v.copy() if isinstance(self.v, np.ndarray) else self.v.clone()
        return HPCVector(self.structural_hash, self.partition, v, self.backend)

    def conj(self) -> "HPCVector":
        """conj(v) — src/vectors.jl:729-733 (elementwise; library-level op, not a hot-path kernel)."""
        v = np.conj(self.v) if isinstance(self.v, np.ndarray) else self.v.conj().resolve_conj()
        return HPCVector(self.structural_hash, self.partition, v, self.backend)

    def __repr__(self):
        return f"HPCVector(n={len(self)}, local={self.local_size}, T={self.backend.T}, {self.backend.device})"


# ---------------------------------------------------------------------------------------------------------------
# reductions / updates on the device (src/vectors.jl:758-812, 1203-1221)
# ---------------------------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------------
# repartition (src/vectors.jl:469-720): the step in front of the hot path whenever partitions differ
# ---------------------------------------------------------------------------------------------------------------
class VectorRepartitionPlan:
    """VectorRepartitionPlan{T} (src/vectors.jl:491-506): the index fields (1-based, ranks ascending), computed by
    hpcla_repartition_plan.  No buffers: the device exchange moves contiguous ranges in place."""

    def __init__(self, rank: int, nranks: int, old_partition: np.ndarray, new_partition: np.ndarray):
        op = np.ascontiguousarray(old_partition, dtype=np.int64)
        npart = np.ascontiguousarray(new_partition, dtype=np.int64)
        if len(npart) != nranks + 1 or len(op) != nranks + 1:
            raise ValueError(f"repartition: a partition must have nranks+1 = {nranks + 1} entries")
        a = [np.zeros(max(nranks, 1), dtype=np.int64) for _ in range(6)]
        ns, nr, size = np.zeros(1, np.int64), np.zeros(1, np.int64), np.zeros(1, np.int64)
        self._local3 = np.zeros(3, dtype=np.int64)
        _lib.check(_lib.lib().hpcla_repartition_plan(rank, nranks, _lib.ptr(op), _lib.ptr(npart), _lib.ptr(ns), _lib.ptr(a[0]), _lib.ptr(a[1]), _lib.ptr(a[2]),
                                                     _lib.ptr(nr), _lib.ptr(a[3]), _lib.ptr(a[4]), _lib.ptr(a[5]), _lib.ptr(self._local3), _lib.ptr(size)))
        ns, nr = int(ns[0]), int(nr[0])
        self._send = [np.ascontiguousarray(v[:ns]) for v in a[:3]]
        self._recv = [np.ascontiguousarray(v[:nr]) for v in a[3:]]
        self.send_rank_ids = self._send[0].tolist()
        self.send_ranges = [(int(f), int(f + c - 1)) for f, c in zip(self._send[1], self._send[2])]
        self.recv_rank_ids = self._recv[0].tolist()
        self.recv_counts = self._recv[1].tolist()
        self.recv_offsets = self._recv[2].tolist()
        f, c, o = (int(v) for v in self._local3)
        self.local_src_range = (f, f + c - 1)
        self.local_dst_offset = o
        self.result_partition = npart.copy()
        self.result_partition_hash = compute_partition_hash(npart)
        self.result_local_size = int(size[0])


_repartition_plan_cache = {}
repartition_plan_build_count = 0


def get_repartition_plan(x: HPCVector, p: np.ndarray) -> VectorRepartitionPlan:
    """get_repartition_plan(x, p) — src/vectors.jl:684-693: memoised on (hash of x's partition, hash of p, T)."""
    global repartition_plan_build_count
    b = x.backend
    key = (x.structural_hash, compute_partition_hash(p), np.dtype(b.T).str, comm_rank(b.comm), id(getattr(b.comm, "world", None)))
    plan = _repartition_plan_cache.get(key)
    if plan is None:
        plan = _repartition_plan_cache[key] = VectorRepartitionPlan(comm_rank(b.comm), comm_size(b.comm), x.partition, p)
        repartition_plan_build_count += 1
    return plan


def repartition(x: HPCVector, p) -> HPCVector:
    """repartition(x, p) — src/vectors.jl:712-720.  Same partition: returns x itself (the reference's fast path)."""
    p = np.ascontiguousarray(p, dtype=np.int64)
    if x.partition is p or (len(x.partition) == len(p) and np.array_equal(x.partition, p)):
        return x
    if len(p) != comm_size(x.backend.comm) + 1 or p[0] != 1 or p[-1] != len(x) + 1:
        raise ValueError(f"repartition: p must have nranks+1 entries, start at 1 and end at length(x)+1 = {len(x) + 1}")
    b = x.backend
    plan = get_repartition_plan(x, p)
    if not b.is_cuda:
        raise _lib.HPCLAError("repartition needs DeviceCUDA operands: this build moves data on the device only (no CPU fallback)")
    import torch

    out = torch.empty(plan.result_local_size, dtype=x.v.dtype, device=x.v.device)
    ctx = b.ctx()
    if ctx.world == "threads":
        # single-process harness world: every rank publishes its slice, the receivers copy their ranges device to device
        srcs = comm_allgather(b.comm, x.v)
        f, l = plan.local_src_range
        if l >= f:
            out[plan.local_dst_offset - 1 : plan.local_dst_offset - 1 + (l - f + 1)].copy_(x.v[f - 1 : l])
        me = comm_rank(b.comm)
        for r, cnt, off in zip(plan.recv_rank_ids, plan.recv_counts, plan.recv_offsets):
            peer = VectorRepartitionPlan(r, comm_size(b.comm), x.partition, p)  # where my range starts in r's slice
            pf = peer.send_ranges[peer.send_rank_ids.index(me)][0]
            out[off - 1 : off - 1 + cnt].copy_(srcs[r][pf - 1 : pf - 1 + cnt])
        torch.cuda.synchronize(x.v.device)
        comm_allgather(b.comm, 0)  # nobody drops its slice before every reader is done
    else:
        s, r = plan._send, plan._recv
        _lib.check(_lib.lib().hpcla_repartition_run(ctx.handle, _lib.dtype_code(b.T), len(plan.send_rank_ids), _lib.ptr(s[0]), _lib.ptr(s[1]), _lib.ptr(s[2]),
                                                    len(plan.recv_rank_ids), _lib.ptr(r[0]), _lib.ptr(r[1]), _lib.ptr(r[2]), _lib.ptr(plan._local3),
                                                    _lib.ptr(x.v), _lib.ptr(out), _current_stream(b)))
    return HPCVector(plan.result_partition_hash, plan.result_partition, out, b)


def _aligned(x: HPCVector, y: HPCVector) -> HPCVector:
    """y on x's partition — what the reference does in front of dot and the vector updates when partitions differ
    (src/vectors.jl:806-811, 873-876)."""
    assert_backends_compatible(x.backend, y.backend)
    if len(x) != len(y):
        raise ValueError(f"DimensionMismatch: lengths {len(x)} and {len(y)}")
    return y if x.structural_hash == y.structural_hash else repartition(y, x.partition)


def dot(x: HPCVector, y: HPCVector):
    """LinearAlgebra.dot(x, y) — src/vectors.jl:798-812: local dot (conjugating x) + Allreduce(+); y is repartitioned
    to x's partition first when they differ (:806-811)."""
    y = _aligned(x, y)
    b = x.backend
    ctx = b.ctx()
    code = _lib.dtype_code(b.T)
    out = np.zeros(2, dtype=np.float64)
    res = np.zeros(1, dtype=np.float32) if code == _lib.F32 else out
    _lib.check(_lib.lib().hpcla_dot(ctx.handle, code, x.local_size, _lib.ptr(x.v), _lib.ptr(y.v), _lib.ptr(res), _current_stream(b)))
    val = complex(out[0], out[1]) if code == _lib.C128 else (float(res[0]) if code == _lib.F32 else float(out[0]))
    if ctx.world == "threads":
        val = comm_allreduce(b.comm, val, "+")
    return b.T.type(val)


def norm(x: HPCVector, p=2):
    """LinearAlgebra.norm(v, 2) — src/vectors.jl:758-766: sqrt(Allreduce(+)(local_norm^2))."""
    if p != 2:
        raise NotImplementedError("only the 2-norm is part of the SpMV/CG hot path (SURVEY §2.1 #3)")
    b = x.backend
    ctx = b.ctx()
    code = _lib.dtype_code(b.T)
    res = np.zeros(1, dtype=np.float32 if code == _lib.F32 else np.float64)
    _lib.check(_lib.lib().hpcla_nrm2(ctx.handle, code, x.local_size, _lib.ptr(x.v), _lib.ptr(res), _current_stream(b)))
    val = float(res[0])
    if ctx.world == "threads":  # the library returned the local sum of squares
        val = float(np.sqrt(comm_allreduce(b.comm, val, "+")))
    return val


def axpby(alpha, x: HPCVector, beta, y: HPCVector) -> HPCVector:
    """y .= alpha .* x .+ beta .* y (the fused broadcast of src/vectors.jl:1203-1221), in place on y; x is
    repartitioned to y's partition first when they differ."""
    x = _aligned(y, x)
    b = x.backend
    a = np.array([alpha], dtype=b.T)
    c = np.array([beta], dtype=b.T)
    _lib.check(_lib.lib().hpcla_axpby(b.ctx().handle, _lib.dtype_code(b.T), x.local_size, _lib.ptr(a), _lib.ptr(x.v), _lib.ptr(c), _lib.ptr(y.v), _current_stream(b)))
    return y
