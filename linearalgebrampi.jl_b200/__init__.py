"""hpcla_b200 — B200-native device backend for the distributed SpMV hot path of HPCLinearAlgebra.jl
(sloisel/LinearAlgebraMPI.jl): `mul!(y, A, x)`, `A * x`, `transpose(A) * x` on a row-partitioned HPCSparseMatrix
times an HPCVector, with the memoised ghost-gather plan.  The public names are the reference's
(src/HPCLinearAlgebra.jl:12-38); the arithmetic and the halo exchange run in libhpcla_b200.so (hand-written sm_100a
CUDA + NCCL).  Import as `import hpcla_b200` (the loader module at the repository root) — the directory name
contains a dot and cannot be imported by name.
"""
from ._lib import HPCLAError, LIB_PATH  # noqa: F401
from .backends import (  # noqa: F401
    AbstractComm, AbstractDevice, AbstractSolver, CommMPI, CommSerial, CommThreads, DeviceCPU, DeviceCUDA, HPCBackend,
    SolverCuDSS, SolverMUMPS, ThreadWorld, assert_backends_compatible, backend_cpu_mpi, backend_cpu_serial,
    backend_cuda_mpi, backend_cuda_serial, backends_compatible, backends_threads, comm_allgather, comm_allreduce,
    comm_barrier, comm_bcast, comm_exchange, comm_rank, comm_size, eltype_backend, indextype_backend, retype_backend,
)
from .vectors import (  # noqa: F401
    HPCVector, VectorRepartitionPlan, axpby, compute_partition_hash, dot, get_repartition_plan, norm, repartition, uniform_partition,
)
from .sparse import (  # noqa: F401
    HPCSparseMatrix, Transpose, VectorPlan, build_vector_plan, cache_sizes, cg, clear_plan_cache, compute_structural_hash, enable_direct_halo,
    execute_plan, get_vector_plan, host_buffer, materialize_transpose, matvec, mul, mul_graph, mul_staged, spmv_info, spmv_timeline, to_backend, transpose, transpose_matvec,
    vec_adjoint_mul, vec_transpose_mul,
)
from .dense import HPCMatrix, spmm  # noqa: F401
from .spgemm import MatrixPlan, get_matrix_plan, spgemm  # noqa: F401
from . import sparse, synth, vectors, backends, dense  # noqa: F401
import importlib as _importlib

spgemm_module = _importlib.import_module(__name__ + ".spgemm")  # the submodule (the name `spgemm` is the function)

HPCVector_local = HPCVector.from_local
HPCMatrix_local = HPCMatrix.from_local
HPCSparseMatrix_local = HPCSparseMatrix.from_local
