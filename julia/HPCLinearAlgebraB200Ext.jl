# HPCLinearAlgebraB200Ext.jl — the reference-side binding of libhpcla_b200.so.
#
# STATUS: EXPERIMENTAL — written against the reference's sources (file:line cited below) but NEVER EXECUTED: this image
# has no julia, no MPI and no CUDA.jl.  The C-ABI route it takes (hpcla_ctx_adopt_nccl / hpcla_ctx_init_nccl,
# hpcla_csr_create, hpcla_plan_import with the reference-built plan, hpcla_spmv_create / run) is exercised call for call
# by tests/test_gpu_round2.py::test_plan_import_route and tests/_nccl_worker.py (section 4) from Python.  It is the stub a maintainer would add next to ext/HPCLinearAlgebraCUDAExt.jl;
# INTEGRATION.md walks through it.  The python mirror under linearalgebrampi.jl_b200/ drives the very same C ABI
# and is what the tests and benchmarks run.
#
# What it does: for DeviceCUDA backends it adds MORE SPECIFIC methods of the hot-path operators, so that
#   A * x, mul!(y, A, x), transpose(A) * x, dot(x, y), norm(x)
# run in hand-written sm_100a kernels with an NCCL halo exchange instead of the CPU-staged MPI path of
# src/vectors.jl:394-463 and the one-thread-per-row KernelAbstractions kernel of src/sparse.jl:2055-2084.
# Everything else of the package (constructors, plan construction with MPI, caches, the CPU backends) is untouched.
module HPCLinearAlgebraB200Ext

using HPCLinearAlgebra
using CUDA
using MPI
using LinearAlgebra
using HPCLinearAlgebra: HPCBackend, HPCVector, HPCSparseMatrix, VectorPlan, DeviceCUDA, CommMPI, CommSerial,
                        comm_rank, comm_size, get_vector_plan, _ensure_hash, compute_partition_hash

const libhpcla = get(ENV, "HPCLA_B200_LIB", "libhpcla_b200.so")

const CuB{T,Ti,C,S} = HPCBackend{T,Ti,DeviceCUDA,C,S}

_dtype(::Type{Float32}) = Cint(0)      # HPCLA_F32
_dtype(::Type{Float64}) = Cint(1)      # HPCLA_F64
_dtype(::Type{ComplexF64}) = Cint(2)   # HPCLA_C128
_itype(::Type{Int32}) = Cint(0)        # HPCLA_I32
_itype(::Type{Int64}) = Cint(1)        # HPCLA_I64

function _check(status::Cint, what::AbstractString)
    # same convention as the cuDSS / NCCL wrappers of ext/HPCLinearAlgebraCUDAExt.jl:247-251, 388-402
    status == 0 || error("$what failed with status $status: " * unsafe_string(@ccall libhpcla.hpcla_last_error()::Cstring))
    return nothing
end

_dptr(a::CuArray) = reinterpret(Ptr{Cvoid}, pointer(a))          # cf. ext/HPCLinearAlgebraCUDAExt.jl:665-669
_stream() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)         # run on the caller's task-local stream

# ---------------------------------------------------------------------------------------------------------------
# context: one per (MPI communicator, device); the NCCL bootstrap is the reference's own
# (ext/HPCLinearAlgebraCUDAExt.jl:411-443): rank 0 creates the id, MPI.Bcast! ships the 128 bytes.
# Never destroyed (ext:384-386: destroying NCCL communicators from finalizers desynchronises ranks).
# ---------------------------------------------------------------------------------------------------------------
const _contexts = Dict{Any,Ptr{Cvoid}}()

function _context(backend::CuB)
    comm = backend.comm
    key = comm isa CommMPI ? comm.comm.val : :serial
    get!(_contexts, key) do
        rank, nranks = comm_rank(comm), comm_size(comm)
        dev = rank % length(CUDA.devices())                        # ext:611-613
        CUDA.device!(dev)
        ctx = Ref{Ptr{Cvoid}}(C_NULL)
        _check(@ccall(libhpcla.hpcla_ctx_create(dev::Cint, rank::Cint, nranks::Cint, ctx::Ptr{Ptr{Cvoid}})::Cint), "hpcla_ctx_create")
        if comm isa CommMPI && nranks > 1
            id = zeros(UInt8, 128)
            rank == 0 && _check(@ccall(libhpcla.hpcla_nccl_unique_id(id::Ptr{UInt8})::Cint), "hpcla_nccl_unique_id")
            MPI.Bcast!(id, 0, comm.comm)
            _check(@ccall(libhpcla.hpcla_ctx_init_nccl(ctx[]::Ptr{Cvoid}, id::Ptr{UInt8})::Cint), "hpcla_ctx_init_nccl")
        end
        ctx[]
    end
end

# ---------------------------------------------------------------------------------------------------------------
# side cache of device-derived state (HPCSparseMatrix has no spare field and HPCVector is immutable, SURVEY §8b):
# keyed like the reference's plan cache (src/sparse.jl:1994) plus the identity of A.nzval; wiped with the plans.
# ---------------------------------------------------------------------------------------------------------------
mutable struct BoundOp
    csr::Ptr{Cvoid}
    plan::Ptr{Cvoid}
    op::Ptr{Cvoid}
    key::Any
end
# Keyed by the matrix OBJECT (weakly): the handles borrow A.rowptr_target / A.colval_target / A.nzval, which live exactly
# as long as A does, so nothing is rooted here and nothing outlives its matrix (the previous version kept every multiplied
# matrix's arrays alive until clear_bound!).  The inner key is the reference's plan-cache key (src/sparse.jl:1994): a
# structural change drops A.structural_hash (src/indexing.jl:1291-1294), hence yields a new key and a new handle.
const _bound = WeakKeyDict{Any,Dict{Any,BoundOp}}()

# Frees the library handles of one matrix.  Only local work (cudaFree + stream synchronisation), no collective — safe in a
# finalizer, unlike destroying an NCCL communicator (ext/HPCLinearAlgebraCUDAExt.jl:9-10).
function _release!(ops::Dict{Any,BoundOp})
    for b in values(ops)
        @ccall libhpcla.hpcla_spmv_destroy(b.op::Ptr{Cvoid})::Cvoid
        @ccall libhpcla.hpcla_plan_destroy(b.plan::Ptr{Cvoid})::Cvoid
        @ccall libhpcla.hpcla_csr_destroy(b.csr::Ptr{Cvoid})::Cvoid
    end
    empty!(ops)
end

function _bind(A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T}, plan::VectorPlan{T,Ti}) where {T,Ti,B<:CuB}
    key = (_ensure_hash(A), x.structural_hash, T, Ti)
    ops = get!(_bound, A) do
        d = Dict{Any,BoundOp}()
        finalizer(_ -> _release!(d), A)   # HPCSparseMatrix is a mutable struct (src/sparse.jl:319): it can carry a finalizer
        d
    end
    get!(ops, key) do
        ctx = _context(A.backend)
        nnz = Int64(length(A.nzval))
        csr = Ref{Ptr{Cvoid}}(C_NULL)
        _check(@ccall(libhpcla.hpcla_csr_create(ctx::Ptr{Cvoid}, _dtype(T)::Cint, _itype(Ti)::Cint, Int64(A.nrows_local)::Int64,
                  Int64(A.ncols_compressed)::Int64, nnz::Int64, _dptr(A.rowptr_target)::Ptr{Cvoid}, _dptr(A.colval_target)::Ptr{Cvoid},
                  _dptr(A.nzval)::Ptr{Cvoid}, csr::Ptr{Ptr{Cvoid}})::Cint), "hpcla_csr_create")
        # hand over the VectorPlan the reference already built with MPI (src/sparse.jl:1875-1984): no second protocol
        # pointer tables as Vector{Ptr{Cvoid}} (the element type the C side declares: const void* const*), taken under
        # GC.@preserve so that the index vectors cannot move or be collected while the library copies them
        send_indices, recv_perm = plan.send_indices, plan.recv_perm
        slen = Int64[length(v) for v in send_indices]
        rlen = Int64[length(v) for v in recv_perm]
        send_ids, recv_ids = Int64.(plan.send_rank_ids), Int64.(plan.recv_rank_ids)
        ph = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve plan send_indices recv_perm send_ids recv_ids begin
            sidx = Ptr{Cvoid}[Ptr{Cvoid}(pointer(v)) for v in send_indices]
            rprm = Ptr{Cvoid}[Ptr{Cvoid}(pointer(v)) for v in recv_perm]
            _check(@ccall(libhpcla.hpcla_plan_import(comm_rank(A.backend.comm)::Cint, comm_size(A.backend.comm)::Cint, _itype(Ti)::Cint,
                      Int64(length(plan.gathered))::Int64, Int64(length(x.v))::Int64,
                      Int64(length(send_ids))::Int64, send_ids::Ptr{Int64}, slen::Ptr{Int64}, sidx::Ptr{Ptr{Cvoid}},
                      Int64(length(recv_ids))::Int64, recv_ids::Ptr{Int64}, rlen::Ptr{Int64}, rprm::Ptr{Ptr{Cvoid}},
                      Int64(length(plan.local_src_indices))::Int64, plan.local_src_indices::Ptr{Cvoid}, plan.local_dst_indices::Ptr{Cvoid},
                      ph::Ptr{Ptr{Cvoid}})::Cint), "hpcla_plan_import")
        end
        op = Ref{Ptr{Cvoid}}(C_NULL)
        _check(@ccall(libhpcla.hpcla_spmv_create(ctx::Ptr{Cvoid}, csr[]::Ptr{Cvoid}, ph[]::Ptr{Cvoid}, Int64(length(x.v))::Int64,
                  op::Ptr{Ptr{Cvoid}})::Cint), "hpcla_spmv_create")
        BoundOp(csr[], ph[], op[], key)
    end
end

function _multiply!(y_local::CuVector{T}, A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T}) where {T,Ti,B<:CuB}
    plan = get_vector_plan(A, x)                                   # memoised, src/sparse.jl:1992-2001
    b = _bind(A, x, plan)
    GC.@preserve x y_local begin
        _check(@ccall(libhpcla.hpcla_spmv_run(b.op::Ptr{Cvoid}, _dptr(x.v)::Ptr{Cvoid}, _dptr(y_local)::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint),
               "hpcla_spmv_run")
    end
    return plan
end

# --- Base.:*(A, x): replaces src/sparse.jl:2096-2128 for CUDA backends -------------------------------------------
function Base.:*(A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T,B}) where {T,Ti,B<:CuB}
    y_local = CUDA.zeros(T, A.nrows_local)
    plan = _multiply!(y_local, A, x)
    if plan.result_partition_hash === nothing                       # src/sparse.jl:2103-2106
        plan.result_partition_hash = compute_partition_hash(A.row_partition)
        plan.result_partition = copy(A.row_partition)
    end
    return HPCVector{T,B}(plan.result_partition_hash, plan.result_partition, y_local, A.backend)
end

# --- mul!(y, A, x): replaces src/sparse.jl:2019-2037 (which multiplies on the CPU) ---------------------------------
function LinearAlgebra.mul!(y::HPCVector{T,B}, A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T,B}) where {T,Ti,B<:CuB}
    _multiply!(y.v, A, x)
    return y
end

# --- staged multiply for host-resident vectors: x_host -> x.v, y.v = A*x, y.v -> y_host, pipelined over row blocks -----
#     (what every A*x of the reference's CUDA path does with two full PCIe copies, src/vectors.jl:423, 460).
#     x_host / y_host: this rank's local slices; pin them (CUDA.pin) for asynchronous copies.  y_host is complete
#     after CUDA.synchronize() of the task-local stream.
function mul_staged!(y_host::Vector{T}, y::HPCVector{T,B}, A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T,B}, x_host::Vector{T}) where {T,Ti,B<:CuB}
    length(x_host) == length(x.v) && length(y_host) == length(y.v) || throw(DimensionMismatch("host slices must match the local slices"))
    plan = get_vector_plan(A, x)
    b = _bind(A, x, plan)
    GC.@preserve x y x_host y_host begin
        _check(@ccall(libhpcla.hpcla_spmv_run_staged(b.op::Ptr{Cvoid}, pointer(x_host)::Ptr{Cvoid}, _dptr(x.v)::Ptr{Cvoid}, _dptr(y.v)::Ptr{Cvoid},
                  pointer(y_host)::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint), "hpcla_spmv_run_staged")
    end
    return y_host
end

# --- A * B::HPCMatrix: replaces the column-by-column loop of src/sparse.jl:2391-2413 (ncols SpMVs, each with a column
#     extraction and its own ghost exchange) by one exchange for all columns and tiles staged once per 4 columns.
function Base.:*(A::HPCSparseMatrix{T,Ti,B}, Bm::HPCMatrix{T,B}) where {T,Ti,B<:CuB}
    ncols = size(Bm, 2)
    col1 = HPCVector_local(view(Bm.A, :, 1), A.backend)               # a column's partition = B's row partition (src/indexing.jl:385-393)
    plan = get_vector_plan(A, col1)
    b = _bind(A, col1, plan)
    C = CUDA.zeros(T, A.nrows_local, ncols)                           # column-major, like B.A
    GC.@preserve Bm C begin
        status = @ccall libhpcla.hpcla_spmm_run(b.op::Ptr{Cvoid}, _dptr(Bm.A)::Ptr{Cvoid}, Int64(stride(Bm.A, 2))::Int64, _dptr(C)::Ptr{Cvoid},
                                                Int64(max(A.nrows_local, 1))::Int64, Cint(ncols)::Cint, _stream()::Ptr{Cvoid})::Cint
        if status == 4   # HPCLA_ERR_STATE: own columns not contiguous in B's rows on some rank -> the reference's loop (all ranks agree:
            return invoke(Base.:*, Tuple{HPCSparseMatrix{T,Ti},HPCMatrix{T}}, A, Bm)   # decide it once, collectively, as the Python mirror does)
        end
        _check(status, "hpcla_spmm_run")
    end
    return HPCMatrix_local(C, A.backend)
end

# --- A * B::HPCSparseMatrix: replaces the host product of src/sparse.jl:991-1059.  The reference's MatrixPlan (structure of
#     B[A.col_indices, :], memoised, src/sparse.jl:579-916) stays as it is; the symbolic product is memoised next to it and
#     every later product is: values of the gathered rows (execute_plan! into a device target, :922-983), one kernel.
const _spgemm = Dict{Any,Ptr{Cvoid}}()
function Base.:*(A::HPCSparseMatrix{T,Ti,B}, Bm::HPCSparseMatrix{T,Ti,B}) where {T,Ti,B<:CuB}
    plan = MatrixPlan(A, Bm)                                          # memoised by the reference
    key = (_ensure_hash(A), _ensure_hash(Bm), T, Ti)
    h = get!(_spgemm, key) do
        AT = plan.AT                                                  # CSC of B[A.col_indices, :]^T: colptr = row pointers, rowval = GLOBAL columns
        out = Ref{Ptr{Cvoid}}(C_NULL)
        _check(@ccall(libhpcla.hpcla_spgemm_symbolic(_itype(Ti)::Cint, Int64(A.nrows_local)::Int64, A.rowptr::Ptr{Cvoid}, A.colval::Ptr{Cvoid},
                  Int64(length(A.col_indices))::Int64, Int64.(AT.colptr)::Ptr{Int64}, Int64.(AT.rowval)::Ptr{Int64}, out::Ptr{Ptr{Cvoid}})::Cint),
               "hpcla_spgemm_symbolic")
        out[]
    end
    nnz = Ref{Int64}(0); ncc = Ref{Int64}(0); nt = Ref{Int64}(0)
    _check(@ccall(libhpcla.hpcla_spgemm_sizes(h::Ptr{Cvoid}, nnz::Ptr{Int64}, ncc::Ptr{Int64}, nt::Ptr{Int64})::Cint), "hpcla_spgemm_sizes")
    rowptr = zeros(Ti, A.nrows_local + 1); colval = zeros(Ti, nnz[]); col_indices = zeros(Int, ncc[])
    _check(@ccall(libhpcla.hpcla_spgemm_structure(h::Ptr{Cvoid}, _itype(Ti)::Cint, rowptr::Ptr{Cvoid}, colval::Ptr{Cvoid}, col_indices::Ptr{Int64})::Cint),
           "hpcla_spgemm_structure")
    bg = CUDA.zeros(T, length(plan.AT.nzval))
    execute_plan!(plan, Bm, bg)                                       # the reference's value gather with a device target (src/sparse.jl:922-983)
    nzval = CUDA.zeros(T, nnz[])
    GC.@preserve A bg nzval begin
        _check(@ccall(libhpcla.hpcla_spgemm_numeric(h::Ptr{Cvoid}, _context(A.backend)::Ptr{Cvoid}, _dtype(T)::Cint, _dptr(A.nzval)::Ptr{Cvoid},
                  _dptr(bg)::Ptr{Cvoid}, _dptr(nzval)::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint), "hpcla_spgemm_numeric")
    end
    return HPCSparseMatrix{T,Ti,B}(nothing, A.row_partition, Bm.col_partition, col_indices, rowptr, colval, nzval, A.nrows_local, ncc[], nothing, nothing,
                                   CuVector(rowptr), CuVector(colval), A.backend)              # as src/sparse.jl:1054-1058
end

# --- repartition(x, p): replaces VectorRepartitionPlan + execute_plan! (src/vectors.jl:519-676: host-staged, tag 92) ------------
function HPCLinearAlgebra.repartition(x::HPCVector{T,B}, p::Vector{Int}) where {T,B<:CuB}
    (x.partition === p || x.partition == p) && return x              # the reference's fast path (src/vectors.jl:714-716)
    nr = length(p) - 1
    rank = comm_rank(x.backend.comm)
    ns = Ref{Int64}(0); nrv = Ref{Int64}(0); sz = Ref{Int64}(0)
    s1 = zeros(Int64, nr); s2 = zeros(Int64, nr); s3 = zeros(Int64, nr)
    r1 = zeros(Int64, nr); r2 = zeros(Int64, nr); r3 = zeros(Int64, nr); loc = zeros(Int64, 3)
    _check(@ccall(libhpcla.hpcla_repartition_plan(Cint(rank)::Cint, Cint(nr)::Cint, x.partition::Ptr{Int64}, p::Ptr{Int64}, ns::Ptr{Int64}, s1::Ptr{Int64},
              s2::Ptr{Int64}, s3::Ptr{Int64}, nrv::Ptr{Int64}, r1::Ptr{Int64}, r2::Ptr{Int64}, r3::Ptr{Int64}, loc::Ptr{Int64}, sz::Ptr{Int64})::Cint),
           "hpcla_repartition_plan")
    out = CUDA.zeros(T, sz[])
    GC.@preserve x out begin
        _check(@ccall(libhpcla.hpcla_repartition_run(_context(x.backend)::Ptr{Cvoid}, _dtype(T)::Cint, ns[]::Int64, s1::Ptr{Int64}, s2::Ptr{Int64}, s3::Ptr{Int64},
                  nrv[]::Int64, r1::Ptr{Int64}, r2::Ptr{Int64}, r3::Ptr{Int64}, loc::Ptr{Int64}, _dptr(x.v)::Ptr{Cvoid}, _dptr(out)::Ptr{Cvoid},
                  _stream()::Ptr{Cvoid})::Cint), "hpcla_repartition_run")
    end
    return HPCVector{T,B}(compute_partition_hash(p), copy(p), out, x.backend)
end

# --- HPCSparseMatrix(transpose(A)): replaces TransposePlan + execute_plan! (src/sparse.jl:1551-1829) and keeps the
#     bidirectional cache of src/sparse.jl:1846-1865
function HPCLinearAlgebra.HPCSparseMatrix(At::Transpose{T,<:HPCSparseMatrix{T,Ti,B}}) where {T,Ti,B<:CuB}
    A = At.parent
    A.cached_transpose !== nothing && return A.cached_transpose
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(@ccall(libhpcla.hpcla_transpose_device(_context(A.backend)::Ptr{Cvoid}, _dtype(T)::Cint, _itype(Ti)::Cint, A.row_partition::Ptr{Int64},
              A.col_partition::Ptr{Int64}, Int64(A.nrows_local)::Int64, Int64(A.ncols_compressed)::Int64, Int64(length(A.nzval))::Int64,
              _dptr(A.rowptr_target)::Ptr{Cvoid}, _dptr(A.colval_target)::Ptr{Cvoid}, A.col_indices::Ptr{Int64}, _dptr(A.nzval)::Ptr{Cvoid},
              _stream()::Ptr{Cvoid}, h::Ptr{Ptr{Cvoid}})::Cint), "hpcla_transpose_device")
    nrows = Ref{Int64}(0); nnz = Ref{Int64}(0); ncc = Ref{Int64}(0)
    _check(@ccall(libhpcla.hpcla_dtb_sizes(h[]::Ptr{Cvoid}, nrows::Ptr{Int64}, nnz::Ptr{Int64}, ncc::Ptr{Int64})::Cint), "hpcla_dtb_sizes")
    rowptr_d = CUDA.zeros(Ti, nrows[] + 1); colval_d = CUDA.zeros(Ti, nnz[]); nzval_d = CUDA.zeros(T, nnz[]); col_indices = zeros(Int, ncc[])
    _check(@ccall(libhpcla.hpcla_dtb_result(h[]::Ptr{Cvoid}, _dptr(rowptr_d)::Ptr{Cvoid}, _dptr(colval_d)::Ptr{Cvoid}, col_indices::Ptr{Int64},
              _dptr(nzval_d)::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint), "hpcla_dtb_result")
    @ccall libhpcla.hpcla_dtb_destroy(h[]::Ptr{Cvoid})::Cvoid
    Y = HPCSparseMatrix{T,Ti,B}(nothing, copy(A.col_partition), copy(A.row_partition), col_indices, Array(rowptr_d), Array(colval_d), nzval_d,
                                Int(nrows[]), Int(ncc[]), nothing, nothing, rowptr_d, colval_d, A.backend)   # field order: src/sparse.jl:319-337
    A.cached_transpose = Y; Y.cached_transpose = A
    return Y
end

# --- transpose(A) * x: src/sparse.jl:2375-2379 already materialises and caches A^T and calls A_transposed * x, which
#     now dispatches to the method above; nothing to override.  (The one-time TransposePlan stays the reference's.)

# --- execute_plan!: replaces src/vectors.jl:394-463 when somebody asks for `gathered` itself ----------------------
#     (needs the matrix the plan belongs to; exposed as a helper rather than an override of the 2-argument method)
function gather!(A::HPCSparseMatrix{T,Ti,B}, x::HPCVector{T,B}) where {T,Ti,B<:CuB}
    plan = get_vector_plan(A, x)
    b = _bind(A, x, plan)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _check(@ccall(libhpcla.hpcla_spmv_gather(b.op::Ptr{Cvoid}, _dptr(x.v)::Ptr{Cvoid}, _stream()::Ptr{Cvoid}, out::Ptr{Ptr{Cvoid}})::Cint), "hpcla_spmv_gather")
    return unsafe_wrap(CuArray, reinterpret(CuPtr{T}, out[]), length(plan.gathered))
end

# --- dot / norm: replace src/vectors.jl:798-812, 758-766 (CUBLAS + MPI.Allreduce of a host scalar) -----------------
function LinearAlgebra.dot(x::HPCVector{T,B}, y::HPCVector{T,B}) where {T,B<:CuB}
    x.structural_hash == y.structural_hash || return invoke(LinearAlgebra.dot, Tuple{HPCVector{T},HPCVector{T}}, x, y)
    r = Ref{T}()
    _check(@ccall(libhpcla.hpcla_dot(_context(x.backend)::Ptr{Cvoid}, _dtype(T)::Cint, Int64(length(x.v))::Int64, _dptr(x.v)::Ptr{Cvoid},
              _dptr(y.v)::Ptr{Cvoid}, r::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint), "hpcla_dot")
    return r[]
end

function LinearAlgebra.norm(x::HPCVector{T,B}, p::Real=2) where {T,B<:CuB}
    p == 2 || return invoke(LinearAlgebra.norm, Tuple{HPCVector{T},Real}, x, p)
    r = Ref{real(T)}()
    _check(@ccall(libhpcla.hpcla_nrm2(_context(x.backend)::Ptr{Cvoid}, _dtype(T)::Cint, Int64(length(x.v))::Int64, _dptr(x.v)::Ptr{Cvoid},
              r::Ptr{Cvoid}, _stream()::Ptr{Cvoid})::Cint), "hpcla_nrm2")
    return r[]
end

# --- cache hygiene: clear_plan_cache!() (src/HPCLinearAlgebra.jl:181-201) must also drop the device state ----------
function clear_bound!()
    for ops in values(_bound)
        _release!(ops)
    end
    return nothing
end

end # module
