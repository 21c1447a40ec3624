"""HPCBackend(Device, Comm, Solver) — host-side mirror of src/backends.jl for the SpMV hot path.

Same names and meaning as the reference:
  * devices  `DeviceCPU`, `DeviceCUDA`                       (src/backends.jl:22-45)
  * comms    `CommSerial`, `CommMPI`                          (src/backends.jl:63, 73-75)
  * solvers  `SolverMUMPS`, `SolverCuDSS` (type tags only: direct solvers are out of scope)
  * `HPCBackend{T,Ti,D,C,S}`                                  (src/backends.jl:137-141)
  * factories `backend_cpu_serial/mpi`, `backend_cuda_serial/mpi` (src/backends.jl:348-376, ext/HPCLinearAlgebraCUDAExt.jl:98-121)
  * `comm_*` primitives                                       (src/backends.jl:207-327)

There is no Julia and no MPI in this image: `CommMPI` wraps a torch.distributed process group (one process per
GPU), which plays the role of the reference's `MPI.Comm`: host-side collectives for plan construction run over
gloo, the per-multiply halo exchange runs inside libhpcla_b200.so over NCCL.  `CommThreads` is a single-process
world of P rank-threads (the CommSerial idea generalised to P ranks) used when fewer GPUs than ranks are available.

`DeviceCPU` backends carry host structure only (partitions, plans, transposes); arithmetic exists on `DeviceCUDA`
alone — there is no CPU fallback.
"""
from __future__ import annotations

import itertools
import threading
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

from . import _lib


# ---------------------------------------------------------------------------------------------------------------
# devices / solvers  (singleton tags, src/backends.jl:22-101)
# ---------------------------------------------------------------------------------------------------------------
class AbstractDevice:
    def __eq__(self, other):
        return type(self) is type(other) and self.__dict__ == other.__dict__

    def __hash__(self):
        return hash((type(self), tuple(sorted(self.__dict__.items()))))

    def __repr__(self):
        return type(self).__name__ + "()"


class DeviceCPU(AbstractDevice):
    pass


class DeviceCUDA(AbstractDevice):
    """The B200 device.  `index` = CUDA ordinal; the reference picks `rank % ndevices` (ext:611-613)."""

    def __init__(self, index: Optional[int] = None):
        self.index = index

    def __repr__(self):
        return f"DeviceCUDA({self.index})"

    def __eq__(self, other):  # backends_compatible compares device TYPES only (src/backends.jl:446)
        return type(self) is type(other)

    def __hash__(self):
        return hash(type(self))


class AbstractSolver:
    def __repr__(self):
        return type(self).__name__ + "()"


class SolverMUMPS(AbstractSolver):
    pass


class SolverCuDSS(AbstractSolver):
    pass


# ---------------------------------------------------------------------------------------------------------------
# comms
# ---------------------------------------------------------------------------------------------------------------
_comm_uid = itertools.count(1)


class AbstractComm:
    pass


class CommSerial(AbstractComm):
    """Single process, nranks = 1 (src/backends.jl:63)."""

    def __repr__(self):
        return "CommSerial()"


class CommMPI(AbstractComm):
    """A torch.distributed process group in the role of `CommMPI(comm::MPI.Comm)` (src/backends.jl:73-75)."""

    def __init__(self, group=None, nccl_comm: Optional[int] = None):
        """nccl_comm: address of an ncclComm_t created by the caller for this group — the communicator the reference
        already caches per MPI communicator (ext/HPCLinearAlgebraCUDAExt.jl:411-443); adopted, not owned, by the
        library (hpcla_ctx_adopt_nccl).  None: the library bootstraps its own (hpcla_ctx_init_nccl)."""
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("CommMPI needs torch.distributed.init_process_group() first (the reference needs MPI.Init())")
        self.group = group
        self.nccl_comm = nccl_comm
        self.uid = next(_comm_uid)
        self._ctxs = {}
        self._dist = dist
        self._host_group = None
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def host_group(self):
        """gloo group for host-side integer collectives (plan construction), whatever the default backend is."""
        dist = self._dist
        if self._host_group is None:
            if dist.get_backend(self.group) == "gloo":
                self._host_group = self.group if self.group is not None else dist.group.WORLD
            else:
                ranks = dist.get_process_group_ranks(self.group if self.group is not None else dist.group.WORLD)
                self._host_group = dist.new_group(ranks=ranks, backend="gloo")
        return self._host_group

    def __repr__(self):
        return f"CommMPI(rank={self.rank}, size={self.size})"


class ThreadWorld:
    """Shared state of a single-process world of P rank-threads."""

    def __init__(self, nranks: int):
        self.nranks = nranks
        self.uid = next(_comm_uid)
        self.barrier = threading.Barrier(nranks)
        self.slots: List[Any] = [None] * nranks
        self.handles: Dict[Any, list] = {}  # device assignment -> raw context handles, while a group is being formed
        self._ctxs: Dict[Any, Any] = {}  # (rank, device) -> DeviceContext
        self.lock = threading.Lock()

    def run(self, fn, *args):
        """Run fn(rank, *args) on P threads (SPMD, like `mpiexec -n P`); returns the list of results."""
        results: List[Any] = [None] * self.nranks
        errors: List[Any] = [None] * self.nranks

        def body(r):
            try:
                results[r] = fn(r, *args)
            except BaseException as e:  # noqa: BLE001 - re-raised below
                errors[r] = e
                self.barrier.abort()

        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.nranks)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        self.barrier.reset()
        for e in errors:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in errors:
            if e is not None:
                raise e
        return results


class CommThreads(AbstractComm):
    """Rank `rank` of a ThreadWorld."""

    def __init__(self, world: ThreadWorld, rank: int):
        self.world = world
        self.rank = rank
        self.size = world.nranks

    def __repr__(self):
        return f"CommThreads(rank={self.rank}, size={self.size})"


# --- comm primitives (src/backends.jl:207-327) -----------------------------------------------------------------
def comm_rank(comm: AbstractComm) -> int:
    return 0 if isinstance(comm, CommSerial) else comm.rank


def comm_size(comm: AbstractComm) -> int:
    return 1 if isinstance(comm, CommSerial) else comm.size


def comm_barrier(comm: AbstractComm) -> None:
    if isinstance(comm, CommMPI):
        comm._dist.barrier(group=comm.host_group())
    elif isinstance(comm, CommThreads):
        comm.world.barrier.wait()


def comm_allgather(comm: AbstractComm, obj) -> list:
    """Allgather of one small python/numpy object per rank (row counts, digests: src/sparse.jl:477-480, :115)."""
    if isinstance(comm, CommSerial):
        return [obj]
    if isinstance(comm, CommMPI):
        out = [None] * comm.size
        comm._dist.all_gather_object(out, obj, group=comm.host_group())
        return out
    w = comm.world
    w.barrier.wait()
    w.slots[comm.rank] = obj
    w.barrier.wait()
    out = list(w.slots)
    w.barrier.wait()
    return out


def comm_allreduce(comm: AbstractComm, value, op="+"):
    """comm_allreduce of one scalar (src/backends.jl:258-268)."""
    vals = comm_allgather(comm, value)
    if op == "+":
        tot = vals[0]
        for v in vals[1:]:
            tot = tot + v
        return tot
    if op == "max":
        return max(vals)
    if op == "min":
        return min(vals)
    raise ValueError(f"unsupported reduction {op!r}")


def comm_bcast(comm: AbstractComm, obj, root: int = 0):
    return comm_allgather(comm, obj if comm_rank(comm) == root else None)[root]


def comm_exchange(comm: AbstractComm, send: Dict[int, np.ndarray], dtype, tag: int) -> Dict[int, np.ndarray]:
    """The reference's message round: Alltoall of counts, then Isend/Irecv!/Waitall of one array per peer
    (src/sparse.jl:1899-1936 with tag 20; :1581-1624 with tag 10; :1770-1796 with tag 11).
    `send[q]` is the array for rank q (missing or empty = no message); returns what each peer sent to this rank."""
    me, P = comm_rank(comm), comm_size(comm)
    dtype = np.dtype(dtype)
    if isinstance(comm, CommSerial):
        return {}
    if isinstance(comm, CommThreads):
        w = comm.world
        w.barrier.wait()
        w.slots[me] = send
        w.barrier.wait()
        out = {}
        for q in range(P):
            if q != me and w.slots[q] is not None:
                a = w.slots[q].get(me)
                if a is not None and len(a):
                    out[q] = np.array(a, dtype=dtype, copy=True)
        w.barrier.wait()
        return out
    import torch

    dist = comm._dist
    g = comm.host_group()
    counts = np.zeros(P, dtype=np.int64)
    for q, a in send.items():
        if q != me:
            counts[q] = len(a)
    all_counts = np.stack(comm_allgather(comm, counts))  # all_counts[src, dst]
    reqs, bufs, keep = [], {}, []
    granks = dist.get_process_group_ranks(g)
    for q in range(P):
        if q == me:
            continue
        if all_counts[me, q] > 0:
            t = torch.from_numpy(np.ascontiguousarray(send[q], dtype=dtype).view(np.uint8).copy())
            keep.append(t)
            reqs.append(dist.isend(t, granks[q], group=g, tag=tag))
        if all_counts[q, me] > 0:
            t = torch.empty(int(all_counts[q, me]) * dtype.itemsize, dtype=torch.uint8)
            bufs[q] = t
            reqs.append(dist.irecv(t, granks[q], group=g, tag=tag))
    for r in reqs:
        r.wait()
    return {q: t.numpy().view(dtype).copy() for q, t in bufs.items()}


# ---------------------------------------------------------------------------------------------------------------
# HPCBackend
# ---------------------------------------------------------------------------------------------------------------
@dataclass(eq=False)
class HPCBackend:
    """HPCBackend{T,Ti,D,C,S} (src/backends.jl:137-141). T and Ti are numpy dtypes."""

    T: np.dtype
    Ti: np.dtype
    device: AbstractDevice
    comm: AbstractComm
    solver: AbstractSolver
    _ctx: Any = field(default=None, repr=False, compare=False)

    def __post_init__(self):
        self.T = np.dtype(self.T)
        self.Ti = np.dtype(self.Ti)
        _lib.dtype_code(self.T)
        _lib.itype_code(self.Ti)

    # --- device context (lazily created; shared by every backend value with the same comm and device) -----------
    @property
    def is_cuda(self) -> bool:
        return isinstance(self.device, DeviceCUDA)

    def ctx(self):
        if not self.is_cuda:
            raise _lib.HPCLAError("this operation needs a DeviceCUDA backend: the B200 build has no CPU arithmetic")
        if self._ctx is None:
            self._ctx = _device_context(self)
        return self._ctx

    def torch_device(self):
        import torch

        return torch.device("cuda", self.ctx().device) if self.is_cuda else torch.device("cpu")


def eltype_backend(b: HPCBackend):
    return b.T


def indextype_backend(b: HPCBackend):
    return b.Ti


class DeviceContext:
    """Owner of one hpcla_ctx (one rank <-> one GPU)."""

    def __init__(self, handle: int, device: int, rank: int, nranks: int, world: str):
        self.handle = handle
        self.device = device
        self.rank = rank
        self.nranks = nranks
        self.world = world  # "single" | "nccl" | "threads"

    def sync(self):
        _lib.check(_lib.lib().hpcla_ctx_sync(self.handle))


_serial_ctx_cache: Dict[int, DeviceContext] = {}  # CommSerial: one context per device
_ctx_lock = threading.Lock()


def comm_uid(comm: AbstractComm) -> int:
    """Identity of a communicator that is never reused (unlike id()); part of the plan-cache key."""
    if isinstance(comm, CommSerial):
        return 0
    return comm.world.uid if isinstance(comm, CommThreads) else comm.uid


def _device_context(b: HPCBackend) -> DeviceContext:
    import ctypes

    import torch

    L = _lib.lib()
    comm = b.comm
    rank, size = comm_rank(comm), comm_size(comm)
    if not torch.cuda.is_available():
        raise _lib.HPCLAError("no CUDA device is visible: DeviceCUDA backends cannot run here (no CPU fallback)")
    ndev = torch.cuda.device_count()
    dev = b.device.index if b.device.index is not None else rank % ndev  # ext/HPCLinearAlgebraCUDAExt.jl:611-613
    # contexts live on the communicator they belong to (never in a cache keyed by id(): ids are reused)
    if isinstance(comm, CommSerial):
        cache, key = _serial_ctx_cache, dev
    elif isinstance(comm, CommThreads):
        cache, key = comm.world._ctxs, (rank, dev)
    else:
        cache, key = comm._ctxs, dev
    with _ctx_lock:
        if key in cache:
            return cache[key]
    h = ctypes.c_void_p()
    _lib.check(L.hpcla_ctx_create(dev, rank, size, ctypes.byref(h)))
    if isinstance(comm, CommSerial) or size == 1:
        world = "single"
    elif isinstance(comm, CommMPI) and comm.nccl_comm:
        torch.cuda.set_device(dev)
        _lib.check(L.hpcla_ctx_adopt_nccl(h, comm.nccl_comm))  # the caller's cached communicator (ext:411-443)
        world = "nccl"
    elif isinstance(comm, CommMPI):
        # NCCL bootstrap as ext/HPCLinearAlgebraCUDAExt.jl:411-443: rank 0 makes the id, the host comm broadcasts it
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.check(L.hpcla_nccl_unique_id(_lib.ptr(ident)))
        ident = np.frombuffer(comm_bcast(comm, ident.tobytes(), 0), dtype=np.uint8).copy()
        torch.cuda.set_device(dev)
        _lib.check(L.hpcla_ctx_init_nccl(h, _lib.ptr(ident)))
        world = "nccl"
    else:
        w = comm.world
        with w.lock:
            w.handles.setdefault("forming", [None] * size)[rank] = h.value
        w.barrier.wait()
        if rank == 0:
            arr = (ctypes.c_void_p * size)(*w.handles.pop("forming"))
            _lib.check(L.hpcla_ctx_form_group(arr, size))
        w.barrier.wait()
        world = "threads"
    ctx = DeviceContext(h.value, dev, rank, size, world)
    with _ctx_lock:
        cache[key] = ctx
    return ctx


# --- factories (src/backends.jl:348-376; ext/HPCLinearAlgebraCUDAExt.jl:98-121) ----------------------------------
def backend_cpu_serial(T=np.float64, Ti=np.int64) -> HPCBackend:
    return HPCBackend(T, Ti, DeviceCPU(), CommSerial(), SolverMUMPS())


def backend_cpu_mpi(T=np.float64, Ti=np.int64, comm: Optional[CommMPI] = None) -> HPCBackend:
    return HPCBackend(T, Ti, DeviceCPU(), comm if comm is not None else CommMPI(), SolverMUMPS())


def backend_cuda_serial(T=np.float64, Ti=np.int64, device: Optional[int] = None) -> HPCBackend:
    return HPCBackend(T, Ti, DeviceCUDA(device), CommSerial(), SolverCuDSS())


def backend_cuda_mpi(T=np.float64, Ti=np.int64, comm: Optional[CommMPI] = None, device: Optional[int] = None) -> HPCBackend:
    return HPCBackend(T, Ti, DeviceCUDA(device), comm if comm is not None else CommMPI(), SolverCuDSS())


def backends_threads(nranks: int, T=np.float64, Ti=np.int64, cuda: bool = True) -> List[HPCBackend]:
    """One backend per rank-thread of a fresh single-process world (use with `backend.comm.world.run`)."""
    w = ThreadWorld(nranks)
    dev = (lambda: DeviceCUDA(None)) if cuda else DeviceCPU
    return [HPCBackend(T, Ti, dev(), CommThreads(w, r), SolverCuDSS() if cuda else SolverMUMPS()) for r in range(nranks)]


def backends_compatible(b1: HPCBackend, b2: HPCBackend) -> bool:
    """src/backends.jl:444-453: same device type, same comm type, and for CommMPI the same communicator."""
    if type(b1.device) is not type(b2.device):
        return False
    if type(b1.comm) is not type(b2.comm):
        return False
    if isinstance(b1.comm, CommMPI) and b1.comm.group is not b2.comm.group:
        return False
    if isinstance(b1.comm, CommThreads) and b1.comm.world is not b2.comm.world:
        return False
    return True


def assert_backends_compatible(b1: HPCBackend, b2: HPCBackend) -> None:
    if not backends_compatible(b1, b2):
        raise ValueError(f"Incompatible backends: {b1} vs {b2}")  # src/backends.jl:460-464


def retype_backend(b: HPCBackend, Tnew) -> HPCBackend:
    """src/backends.jl:482-487"""
    if np.dtype(Tnew) == b.T:
        return b
    nb = HPCBackend(Tnew, b.Ti, b.device, b.comm, b.solver)
    nb._ctx = b._ctx
    return nb
