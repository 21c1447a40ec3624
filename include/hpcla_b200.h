/* hpcla_b200.h — C ABI of libhpcla_b200.so: the B200 (sm_100a) device backend for the distributed sparse
 * mat-vec hot path of HPCLinearAlgebra.jl (sloisel/LinearAlgebraMPI.jl).
 *
 * Every entry point names the reference interface it replaces (file:line under the reference tree).
 * Conventions:
 *   - plain pointers and sizes only; no C++/torch/Julia types cross this boundary;
 *   - every function returns int, 0 = HPCLA_OK; the message of the last failure on the calling thread is
 *     hpcla_last_error(); nothing here aborts or throws across the ABI (cf. the status checks of
 *     ext/HPCLinearAlgebraCUDAExt.jl:247-251, 388-402);
 *   - index arrays hold the reference's own 1-BASED values, in the reference's own width Ti (Int32 or Int64);
 *     partitions and col_indices are Int64 as in the reference (Vector{Int});
 *   - device arrays are BORROWED for the lifetime of the handle created from them (the caller keeps them rooted);
 *     the library owns and frees only what it allocates itself;
 *   - device work is enqueued on the caller's stream (cudaStream_t passed as void*) and the call returns after
 *     enqueue, like the reference's _spmv! (src/sparse.jl:2081-2083); hpcla_ctx_sync() blocks;
 *   - collective entry points must be called by all ranks in the same order (docs/src/api.md:5-6).
 */
#ifndef HPCLA_B200_H
#define HPCLA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPCLA_ABI_VERSION 1

/* status codes */
enum { HPCLA_OK = 0, HPCLA_ERR_ARG = 1, HPCLA_ERR_CUDA = 2, HPCLA_ERR_NCCL = 3, HPCLA_ERR_STATE = 4, HPCLA_ERR_NOMEM = 5 };
/* element type T of HPCBackend{T,Ti,...} (src/backends.jl:137-141) */
enum { HPCLA_F32 = 0, HPCLA_F64 = 1, HPCLA_C128 = 2 };
/* index type Ti of HPCBackend{T,Ti,...} */
enum { HPCLA_I32 = 0, HPCLA_I64 = 1 };

typedef struct hpcla_ctx hpcla_ctx;     /* one rank <-> one GPU (ext/HPCLinearAlgebraCUDAExt.jl:611-613)          */
typedef struct hpcla_csr hpcla_csr;     /* device view of one rank's HPCSparseMatrix (src/sparse.jl:319-337)       */
typedef struct hpcla_planb hpcla_planb; /* VectorPlan under construction (between the two message rounds)          */
typedef struct hpcla_plan hpcla_plan;   /* index fields of a VectorPlan (src/vectors.jl:229-251)                   */
typedef struct hpcla_spmv hpcla_spmv;   /* (csr, plan) bound to device buffers: what `A*x` / `mul!` execute         */
typedef struct hpcla_tb hpcla_tb;       /* TransposePlan under construction / its result (src/sparse.jl:1519-1538) */
typedef struct hpcla_dtb hpcla_dtb;     /* the same result, built and held on the device                          */
typedef struct hpcla_spgemm hpcla_spgemm; /* memoised symbolic product of A and the gathered rows of B              */

int hpcla_abi_version(void);
const char* hpcla_last_error(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Context.  Replaces the device selection + NCCL bootstrap of ext/HPCLinearAlgebraCUDAExt.jl:370-443, 611-613.
 * ------------------------------------------------------------------------------------------------------------- */
/* Binds `device` (cudaSetDevice) to rank `rank` of `nranks`; creates the halo stream and events. */
int hpcla_ctx_create(int device, int rank, int nranks, hpcla_ctx** out);
/* ncclGetUniqueId (ext:388-393): 128 bytes, produced on rank 0 and broadcast by the host (MPI.Bcast! / torch). */
int hpcla_nccl_unique_id(void* id128);
/* ncclCommInitRank (ext:395-402).  Collective. */
int hpcla_ctx_init_nccl(hpcla_ctx* ctx, const void* id128);
/* Reuse the ncclComm_t the reference already caches per MPI communicator (ext:411-443). Not owned. */
int hpcla_ctx_adopt_nccl(hpcla_ctx* ctx, void* nccl_comm);
/* Single-process world: `n` contexts (ranks 0..n-1, any device assignment) exchange halos with device-to-device
 * copies instead of NCCL.  This is the CommSerial-like harness mode (src/backends.jl:63) generalised to P ranks. */
int hpcla_ctx_form_group(hpcla_ctx* const* ctxs, int n);
int hpcla_ctx_sync(hpcla_ctx* ctx);
void hpcla_ctx_destroy(hpcla_ctx* ctx);
/* Pinned host memory for the staged multiply (hpcla_spmv_run_staged), allocated and first-touched on the CPUs the GPU's
 * PCIe root is attached to (sysfs local_cpulist), so that on a multi-socket host the H2D / D2H copies stay on the GPU's
 * own NUMA node.  The host-side counterpart of the reference's `Array(x.v)` / `copyto!` staging buffers
 * (src/vectors.jl:423, 460), which live wherever Julia's allocator put them.  numa_node_out (may be NULL): the node
 * read from sysfs, or -1.  Falls back to a plain pinned allocation when the topology cannot be read. */
int hpcla_host_alloc(hpcla_ctx* ctx, int64_t bytes, void** out, int* numa_node_out);
int hpcla_host_free(hpcla_ctx* ctx, void* p);

/* ---------------------------------------------------------------------------------------------------------------
 * Host-side structure (pure functions, no device).
 * ------------------------------------------------------------------------------------------------------------- */
/* uniform_partition(n, nranks) — src/HPCLinearAlgebra.jl:279-289. partition_out[nranks+1], 1-based starts. */
int hpcla_uniform_partition(int64_t n, int nranks, int64_t* partition_out);
/* col_indices = unique!(sort(global cols)) and colval = searchsortedfirst(col_indices, g)
 * — src/sparse.jl:501, 504 and compress_AT :137-144.
 * global_cols: Ti[nnz] 1-based global columns; colval_out: Ti[nnz]; col_indices_out: capacity min(nnz, ncols_global). */
int hpcla_compress_columns(int itype, int64_t nnz, const void* global_cols, int64_t ncols_global, void* colval_out,
                           int64_t* col_indices_out, int64_t* ncols_compressed_out);

/* VectorPlan(A, x) — src/sparse.jl:1875-1984 — split at its two message rounds so that the HOST does the
 * communication (MPI in Julia, torch.distributed in the Python harness, plain moves in a single-process world):
 *   begin    = Step 1 (:1886-1895) grouping of col_indices by owner in x.partition
 *   counts   = the send_counts vector handed to the Alltoall (:1899-1900)
 *   requests = the tag-20 message for `owner` (:1908-1918): ascending global indices wanted from it
 *   finish   = Steps 4-7 (:1925-1957) given what the Alltoall and the tag-20 receives delivered. Consumes `pb`. */
int hpcla_plan_begin(int rank, int nranks, const int64_t* col_indices, int64_t ncols_compressed,
                     const int64_t* x_partition, hpcla_planb** out);
int hpcla_planb_counts(const hpcla_planb* pb, int64_t* counts_out /* [nranks] */);
int hpcla_planb_requests(const hpcla_planb* pb, int owner, int64_t* globals_out);
int hpcla_plan_finish(hpcla_planb* pb, const int64_t* recv_counts /* [nranks] */,
                      const int64_t* const* recv_lists /* [nranks], entry r = globals rank r wants from me */,
                      hpcla_plan** out);
/* Import a VectorPlan the reference already built (all fields of src/vectors.jl:229-251 that are indices).
 * Index arrays are Ti (itype), rank ids Int64. n_x_local = length(x.v). */
int hpcla_plan_import(int rank, int nranks, int itype, int64_t n_gathered, int64_t n_x_local, int64_t n_send,
                      const int64_t* send_rank_ids, const int64_t* send_lens, const void* const* send_indices,
                      int64_t n_recv, const int64_t* recv_rank_ids, const int64_t* recv_lens,
                      const void* const* recv_perm, int64_t n_local, const void* local_src_indices,
                      const void* local_dst_indices, hpcla_plan** out);
/* Export, for bit-exact comparison.  field: 0 send_rank_ids, 1 recv_rank_ids, 2 local_src_indices,
 * 3 local_dst_indices, 4 send_indices[slot], 5 recv_perm[slot]; values as Int64. */
int hpcla_plan_len(const hpcla_plan* plan, int field, int64_t slot, int64_t* len_out);
int hpcla_plan_get(const hpcla_plan* plan, int field, int64_t slot, int64_t* out);
int hpcla_plan_n_gathered(const hpcla_plan* plan, int64_t* n_out);
void hpcla_plan_destroy(hpcla_plan* plan);

/* TransposePlan(A) + execute_plan!(plan, A) — src/sparse.jl:1551-1744, 1756-1829 — split at its message round
 * (Alltoall of counts :1581-1583, tag-10 (row,col) pairs :1589-1624, tag-11 values :1770-1796):
 *   begin   = Step 1: bucket every local nonzero by the owner of its column in col_partition
 *   counts  = nonzeros destined to each rank (own rank included)
 *   message = for `dest`: pairs (j = row of A^T, i = col of A^T) in the reference's send order, and the values
 *   finish  = Steps 5-6 + compression: CSR of the owned rows of A^T, ascending columns, compressed, Ti indices
 *   result  = copy out rowptr[nrows+1] (Ti), colval[nnz] (Ti), col_indices[ncc] (Int64), nzval[nnz] (T)        */
int hpcla_transpose_begin(int rank, int nranks, int dtype, int itype, const int64_t* row_partition,
                          const int64_t* col_partition, const void* h_rowptr, const void* h_colval,
                          const int64_t* h_col_indices, const void* h_nzval, hpcla_tb** out);
int hpcla_tb_counts(const hpcla_tb* tb, int64_t* counts_out /* [nranks] */);
int hpcla_tb_message(const hpcla_tb* tb, int dest, int64_t* pairs_out /* [2*count] */, void* vals_out /* T[count] */);
int hpcla_transpose_finish(hpcla_tb* tb, const int64_t* recv_counts, const int64_t* const* recv_pairs,
                           const void* const* recv_vals, int64_t* nrows_out, int64_t* nnz_out, int64_t* ncc_out);
int hpcla_tb_result(const hpcla_tb* tb, void* rowptr_out, void* colval_out, int64_t* col_indices_out, void* nzval_out);
void hpcla_tb_destroy(hpcla_tb* tb);

/* The same materialisation entirely on the device: TransposePlan(A) + execute_plan!(plan, A) —
 * src/sparse.jl:1551-1744, 1756-1829 — without the host copies, the host-staged tag-10/tag-11 messages and the host
 * tuple sort (:1655).  Inputs are A's device arrays as for hpcla_csr_create plus its host partitions and col_indices.
 * Every local nonzero is keyed (row of A^T local to its owner, global column of A^T), dropped into its owner's send
 * range, exchanged with grouped ncclSend/ncclRecv, radix-sorted (CUB), and turned into CSR with compressed columns.
 * Collective (NCCL world or nranks == 1).  Results are bit-identical to hpcla_transpose_begin/finish.
 *   sizes  = rows (of A^T owned here), stored entries, compressed columns
 *   result = device-to-device copies into caller-owned arrays (rowptr Ti[nrows+1], colval Ti[nnz], nzval T[nnz]) and
 *            col_indices (Int64[ncc]) to the host; blocks until done. */
int hpcla_transpose_device(hpcla_ctx* ctx, int dtype, int itype, const int64_t* row_partition,
                           const int64_t* col_partition, int64_t nrows_local, int64_t ncols_compressed, int64_t nnz,
                           const void* d_rowptr, const void* d_colval, const int64_t* h_col_indices,
                           const void* d_nzval, void* stream, hpcla_dtb** out);
int hpcla_dtb_sizes(const hpcla_dtb* t, int64_t* nrows_out, int64_t* nnz_out, int64_t* ncc_out);
int hpcla_dtb_result(const hpcla_dtb* t, void* d_rowptr_out, void* d_colval_out, int64_t* h_col_indices_out,
                     void* d_nzval_out, void* stream);
void hpcla_dtb_destroy(hpcla_dtb* t);

/* ---------------------------------------------------------------------------------------------------------------
 * Device objects.
 * ------------------------------------------------------------------------------------------------------------- */
/* Device view of A: borrows A.rowptr_target / A.colval_target / A.nzval (src/sparse.jl:334-335, 519-520), 1-based
 * contents, unchanged, after one validation pass (rowptr[1] == 1, non-decreasing, rowptr[end] == nnz + 1,
 * 1 <= colval <= ncols_compressed; HPCLA_ERR_ARG otherwise).  A.nzval is always read in place, so in-place value writes —
 * src/indexing.jl:932-982 — are seen by the next multiply.  Library-owned STRUCTURE derived from rowptr / colval is
 * kept beside the matrix: the row-tile table, re-blocked 16-bit row offsets per tile (direct row walk), per (matrix,
 * plan) the x runs and 16-bit column positions of the interior tiles (compact row walk), and for irregular matrices one
 * row-start bit per stored entry (nnz-split kernel).  A structural change needs a new handle, exactly as it drops
 * A.structural_hash and the cached plans in the reference (src/indexing.jl:1291-1294). */
int hpcla_csr_create(hpcla_ctx* ctx, int dtype, int itype, int64_t nrows_local, int64_t ncols_compressed, int64_t nnz,
                     const void* d_rowptr, const void* d_colval, const void* d_nzval, hpcla_csr** out);
/* query: number of row tiles, rows longer than the split threshold, and the lanes per row of the row-walk kernel
 * picked from the mean row length (1, 2, 4, ... 32; 0 = the matrix is irregular and uses the general kernel only) */
int hpcla_csr_info(const hpcla_csr* csr, int64_t* ntiles_out, int64_t* nlong_rows_out, int* lanes_out);
/* query: tiles taken by the TMA-staged row-walk kernel / by the general kernel / holding no row, and the tile window
 * (stored entries per tile) */
int hpcla_csr_tile_classes(const hpcla_csr* csr, int64_t* n_rowwalk_out, int64_t* n_general_out, int64_t* n_empty_out,
                           int* window_out);
void hpcla_csr_destroy(hpcla_csr* csr);

/* Binds (A, VectorPlan) to device buffers: the device copy of send indices, the packed send buffer, `gathered`
 * (src/sparse.jl:1972) and the interior/boundary tile lists.  n_x_local = length(x.v) on this rank.  In a
 * single-process world all ranks must create their operators in the same order (they are paired by sequence). */
int hpcla_spmv_create(hpcla_ctx* ctx, hpcla_csr* csr, const hpcla_plan* plan, int64_t n_x_local, hpcla_spmv** out);
/* y.v = A * x — replaces execute_plan! (src/vectors.jl:394-463) + _spmv! (src/sparse.jl:2075-2084), i.e. the body
 * of Base.:*(A, x) (src/sparse.jl:2096-2128) and of mul!(y, A, x) (:2019-2037).  d_x = x.v, d_y = y.v (device).
 * NCCL world or nranks == 1: one call does everything (collective).  Enqueue only. */
int hpcla_spmv_run(hpcla_spmv* op, const void* d_x, void* d_y, void* stream);
/* Staged multiply for HOST-resident vectors — what the reference's CUDA path does on every call, a host -> device
 * copy of x, the multiply, a device -> host copy of the result (src/vectors.jl:423, 460; mul! computes on host arrays,
 * src/sparse.jl:2019-2037) — as one pipelined call: x.v is uploaded in chunks, each block of rows runs as soon as the
 * prefix of x.v it reads has landed, and its slice of y.v is downloaded while later blocks compute (PCIe both ways at
 * once).  Equivalent to copy(h_x -> d_x); hpcla_spmv_run(op, d_x, d_y); copy(d_y -> h_y).  h_x / h_y should be pinned.
 * Enqueue only: h_y is complete once `stream` has been synchronised.  NCCL world or nranks == 1. */
int hpcla_spmv_run_staged(hpcla_spmv* op, const void* h_x, void* d_x, void* d_y, void* h_y, void* stream);
/* One multiply as a CUDA graph bound to fixed x.v / y.v (NCCL world or a single rank): capture (re)builds it — the
 * event choreography between the caller's stream and the halo stream becomes graph dependencies, the grouped
 * ncclSend/ncclRecv a captured node — and launch replays it with one driver call.  For the latency-bound regime
 * (strong scaling: tens of microseconds of kernel per step against ~10 driver calls). */
int hpcla_spmv_graph_capture(hpcla_spmv* op, const void* d_x, void* d_y, void* stream);
int hpcla_spmv_graph_launch(hpcla_spmv* op, void* stream);
/* Direct halo (optional; grouped ncclSend/ncclRecv stays the default): the ghost runs are PUSHED into the neighbours'
 * `gathered` buffers over NVLink by the copy engine and announced with flags that the receiving halo stream waits on with
 * stream memory operations — no NCCL kernel, no rendezvous, no kernel that spins.  Replaces the same Isend/Irecv/Waitall of
 * src/vectors.jl:431-457.  Collective set-up through the host, like the 128-byte NCCL id (ext:411-443): every rank exports
 * a blob (CUDA IPC handles of its `gathered` and flag buffers + where each neighbour's run lands), the host communicator
 * all-gathers the blobs in rank order, every rank connects.  Works between processes (IPC) and between the rank-threads
 * of one process (plain pointers).  Afterwards hpcla_spmv_run / begin+finish / gather / run_staged / hpcla_cg use it. */
int hpcla_spmv_halo_blob_size(const hpcla_spmv* op, int64_t* bytes_out);
int hpcla_spmv_halo_export(hpcla_spmv* op, void* blob_out);
int hpcla_spmv_halo_connect(hpcla_spmv* op, const void* blobs /* nranks blobs, rank order */);
/* debugging aid: out[0 .. nranks) = last step whose data from each rank has landed, out[nranks .. 2 nranks) = last step each
 * rank has finished reading, out[2 nranks] = this rank's step counter; read on a private stream */
int hpcla_spmv_halo_debug(hpcla_spmv* op, unsigned* out);
/* Timeline of the most recent multiply, for operators created with HPCLA_TIMELINE=1 in the environment: milliseconds
 * from "x ready on the caller's stream" to the end of [0] the halo exchange, [1] the boundary tiles (both on the halo
 * stream), [2] the interior tiles, [3] the whole call (caller's stream); -1 where a step does not exist.  Blocks. */
int hpcla_spmv_timeline(hpcla_spmv* op, double* ms4_out);
/* Single-process world: call begin on every rank, then finish on every rank. */
int hpcla_spmv_begin(hpcla_spmv* op, const void* d_x, void* d_y, void* stream);
int hpcla_spmv_finish(hpcla_spmv* op);
/* C = A * B for a dense right-hand side — replaces Base.:*(A::HPCSparseMatrix, B::HPCMatrix) (src/sparse.jl:2391-2413:
 * a loop of ncols `A * B[:, k]` with a column extraction and a ghost exchange each) by one halo exchange for all
 * columns and kernels that stage every tile of A once per 4 columns.  d_B: B.A, this rank's rows of B, column-major
 * with leading dimension ldb (x.v of column k = d_B + k*ldb, partitioned like the plan's x); d_C: the local block of
 * the result (rows of A.row_partition), column-major, ldc.  Same calling discipline as hpcla_spmv_run / begin+finish.
 * Fails with HPCLA_ERR_STATE when A's own columns are not a contiguous run of B's local rows (the caller then
 * multiplies column by column, like the reference). */
int hpcla_spmm_run(hpcla_spmv* op, const void* d_B, int64_t ldb, void* d_C, int64_t ldc, int ncols, void* stream);
int hpcla_spmm_begin(hpcla_spmv* op, const void* d_B, int64_t ldb, void* d_C, int64_t ldc, int ncols, void* stream);
int hpcla_spmm_finish(hpcla_spmv* op);
/* Parity hook: fills plan.gathered exactly as execute_plan! would and returns its device address
 * (gathered[d] == x_global[col_indices[d]]).  Same calling discipline as run (NCCL) / begin+finish (group: this is
 * the begin half; hpcla_spmv_gather_finish the other). */
int hpcla_spmv_gather(hpcla_spmv* op, const void* d_x, void* stream, void** d_gathered_out);
int hpcla_spmv_gather_finish(hpcla_spmv* op);
/* introspection for tests/benchmarks: counts of interior and boundary tiles, whether x.v is read in place */
int hpcla_spmv_info(const hpcla_spmv* op, int64_t* n_interior_tiles, int64_t* n_boundary_tiles, int* x_in_place,
                    int* sends_contiguous);
/* which kernel takes what: out6 = {row-walk tiles interior, boundary; general tiles interior, boundary; compact row-walk
 * tiles (interior); chunks of the nnz-split kernel (irregular matrices: it then takes every stored entry)} */
int hpcla_spmv_tile_lists(const hpcla_spmv* op, int64_t* out6);
/* kernel launches enqueued by this operator so far (bench.py's gpu_launches) */
int64_t hpcla_spmv_launch_count(const hpcla_spmv* op);
void hpcla_spmv_destroy(hpcla_spmv* op);

/* ---------------------------------------------------------------------------------------------------------------
 * Sparse x sparse, A * B::HPCSparseMatrix — replaces the body of Base.:*(A, B), src/sparse.jl:991-1059,
 * which copies both operands to the host and calls SparseArrays' `plan.AT * A_csc` (symbolic + numeric) on every call.
 *   symbolic  (once per pair of structures, host): A's local CSR (rowptr, colval = 1-based position in A.col_indices,
 *             i.e. the number of the gathered row of B) and the structure of the gathered rows B[A.col_indices, :]
 *             (bg_rowptr: Int64[n+1], 1-based; bg_cols: GLOBAL 1-based columns, ascending within a row — what the
 *             reference's MatrixPlan receives, src/sparse.jl:579-897) -> structure of the local block of C and, per
 *             stored entry of C, the (entry of A, entry of gathered B) pairs that feed it, ascending in the shared index.
 *   sizes     stored entries of C, compressed columns, product terms
 *   structure rowptr Ti[nrows+1], colval Ti[nnz] (compressed), col_indices Int64[ncc] (src/sparse.jl:1018-1040)
 *   numeric   (every call, device): C.nzval[d] = sum of Bg.nzval[ib] * A.nzval[ia] over the pairs of d, in order: the
 *             values of the reference, bit for bit; cancelled entries are kept as stored zeros, as there.
 *             d_bg_nzval: the values of the gathered rows in the order of bg_cols (execute_plan!, src/sparse.jl:922-983). */
int hpcla_spgemm_symbolic(int itype, int64_t nrows_local, const void* a_rowptr, const void* a_colval,
                          int64_t n_gathered_rows, const int64_t* bg_rowptr, const int64_t* bg_cols, hpcla_spgemm** out);
int hpcla_spgemm_sizes(const hpcla_spgemm* plan, int64_t* nnz_out, int64_t* ncc_out, int64_t* nterms_out);
int hpcla_spgemm_structure(const hpcla_spgemm* plan, int itype, void* rowptr_out, void* colval_out, int64_t* col_indices_out);
int hpcla_spgemm_numeric(hpcla_spgemm* plan, hpcla_ctx* ctx, int dtype, const void* d_a_nzval, const void* d_bg_nzval,
                         void* d_c_nzval, void* stream);
void hpcla_spgemm_destroy(hpcla_spgemm* plan);
/* All-to-all of byte ranges between device buffers: rank q gets d_send[send_off[q] .. +send_bytes[q]) into its
 * d_recv[recv_off[me] ..); grouped ncclSend/ncclRecv, the own range is a device-to-device copy.  The value exchange of
 * a MatrixPlan (tag 3 of src/sparse.jl:945-975) without host staging.  Collective (NCCL world or nranks == 1). */
int hpcla_exchange_bytes(hpcla_ctx* ctx, const void* d_send, const int64_t* send_off, const int64_t* send_bytes,
                         void* d_recv, const int64_t* recv_off, const int64_t* recv_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * HPCVector reductions and updates used by iterative solvers.
 * ------------------------------------------------------------------------------------------------------------- */
/* dot(x, y) — src/vectors.jl:798-812: local dot (conjugating x for complex) + Allreduce(+).  result_out: T on the
 * host (2 doubles for C128).  Blocking (the reference returns a host scalar too).  Collective in an NCCL world;
 * in a single-process world it returns the LOCAL dot and the host adds the ranks. */
int hpcla_dot(hpcla_ctx* ctx, int dtype, int64_t n, const void* d_x, const void* d_y, void* result_out, void* stream);
/* norm(v) — src/vectors.jl:758-766: sqrt(Allreduce(+)(local_norm^2)); result_out: double (float for F32).
 * Single-process world: returns the local SUM OF SQUARES instead (host adds and takes the root). */
int hpcla_nrm2(hpcla_ctx* ctx, int dtype, int64_t n, const void* d_x, void* result_out, void* stream);
/* y .= alpha .* x .+ beta .* y — the fused broadcast of src/vectors.jl:1203-1221; alpha, beta: T on the host. */
int hpcla_axpby(hpcla_ctx* ctx, int dtype, int64_t n, const void* alpha, const void* d_x, const void* beta, void* d_y,
                void* stream);
/* VectorRepartitionPlan(x, p) — src/vectors.jl:519-616 — as a pure function of the two partitions (1-based starts,
 * nranks+1 entries each).  Outputs (capacity nranks each): send_rank_ids / send_first (1-based first local index of the
 * range in x.v, = first(send_ranges[i])) / send_count; recv_rank_ids / recv_count / recv_offset (1-based position in
 * the result, = recv_offsets[i]); local3 = {first(local_src_range), length(local_src_range), local_dst_offset};
 * result_local_size.  Same field values as the reference's plan, ranks ascending. */
int hpcla_repartition_plan(int rank, int nranks, const int64_t* old_partition, const int64_t* new_partition,
                           int64_t* n_send_out, int64_t* send_rank_ids, int64_t* send_first, int64_t* send_count,
                           int64_t* n_recv_out, int64_t* recv_rank_ids, int64_t* recv_count, int64_t* recv_offset,
                           int64_t* local3_out, int64_t* result_local_size_out);
/* execute_plan!(plan::VectorRepartitionPlan, x) — src/vectors.jl:624-676 (host-staged Isend/Irecv, tag 92) — on the
 * device: one device-to-device copy for the local overlap, grouped ncclSend/ncclRecv of contiguous ranges for the
 * rest, straight between x.v (d_src) and the result (d_dst, result_local_size elements).  Collective; enqueue only. */
int hpcla_repartition_run(hpcla_ctx* ctx, int dtype, int64_t n_send, const int64_t* send_rank_ids,
                          const int64_t* send_first, const int64_t* send_count, int64_t n_recv,
                          const int64_t* recv_rank_ids, const int64_t* recv_count, const int64_t* recv_offset,
                          const int64_t* local3, const void* d_src, void* d_dst, void* stream);
/* Fixed-iteration conjugate gradients composed from the operations above (the reference ships no CG):
 * x0 = 0, r = b, p = r; iters times { q = A p; alpha = rr/dot(p,q); x += alpha p; r -= alpha q; rr' = dot(r,r);
 * p = r + (rr'/rr) p }.  All scalars stay on the device; no host synchronisation inside the loop.  F32/F64 only.
 * work: device scratch of 3*n elements (r, p, q).  rr_history_out: host double[iters] (may be NULL), filled after a
 * final synchronisation.  NCCL world or nranks == 1. */
int hpcla_cg(hpcla_spmv* op, const void* d_b, void* d_x, void* d_work, int iters, double* rr_history_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPCLA_B200_H */
