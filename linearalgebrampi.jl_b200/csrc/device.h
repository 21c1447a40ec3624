// device.h — internal interface between the orchestration (context.cu) and the kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace hpcla {

// One row tile: rows [row, next.row) whose first stored entry lies in one window of the nonzero stream;
// nnz = 0-based offset of the first stored entry of `row`.
struct TileDesc {
    i64 row;
    i64 nnz;
};

// What one CTA of a multiply kernel needs to know about its tile: rows [r0, r1), 0-based stored entries [s, e).
struct TileRec {
    i64 r0, r1, s, e;
};

// Tile shape of one matrix (see kernels.cu: shape_of).
struct TileShape {
    int threads;        // CTA size of the row-walk kernel
    int lanes;          // row walk: lanes per row (1, 2, 4, ..., 32); 0: the matrix uses the general kernel only
    int window;         // nonzeros per tile window
    int cap;            // row walk: staged nonzeros per tile (window + slack, multiple of 4)
    int rp_cap;         // row walk: staged row pointers per tile (multiple of 4)
    int general_elems;  // general kernel: products staged per tile
    int ovf;            // direct row walk: entries past the window fetched unconditionally
    int hdr_rows;       // direct row walk: rows a tile header can describe (0: no direct walk for this shape)
    int hdr_bytes;      // direct row walk: bytes per tile header (multiple of 16)
};
// avg_row: typical stored entries per row (the most common row length, else the mean); irregular: general kernel only; overrides: 0 = none (tuning hooks)
TileShape tile_shape(int dtype, int itype, double avg_row, bool irregular, int lanes_override, int window_override);
size_t rowwalk_smem_bytes(int dtype, int itype, const TileShape& shape);
size_t direct_smem_bytes(int dtype, int itype, const TileShape& shape);

struct SpmvLaunch {
    int dtype, itype;
    const void* rowptr;
    const void* colval;
    const void* nzval;
    i64 nrows, nnz;
    TileShape shape;        // as fixed when the tile table was built
    const TileRec* recs;    // [n_launch] the tiles of this launch, one CTA each
    // the launch as <= 8 runs of consecutive tiles (run j = CTAs run_cta0[j] .. run_cta0[j+1]-1); n_runs = 0: none
    const unsigned char* hdrs = nullptr;  // direct row walk: tile headers
    int n_runs = 0;
    int run_cta0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int run_tile0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_launch;
    // x addressing for a 1-based compressed column c:
    //   own  <=> own_lo <= c < own_lo + own_n           -> x_own[c - own_lo]   (x_own already offset to the first own source)
    //   else                                            -> gathered[c - 1]
    const void* x_own;
    const void* gathered;
    i64 own_lo, own_n;
    bool has_ghost;  // false: every column is own (single rank, or no off-rank columns)
    void* y;
    i64 long_threshold;  // rows longer than this are left to the split kernels
    // row-walk kernel only: also produce dot(dot_x, y) as one double partial per CTA (dot_x[r] pairs with local row r)
    const void* dot_x = nullptr;
    double* dot_out = nullptr;
};

cudaError_t launch_build_tiles(int itype, const void* rowptr, i64 nrows, i64 nnz, int window, TileDesc* tiles, i64 ntiles,
                               cudaStream_t st);
// cls[t] = 0 (no rows), 1 (row-walk kernel), 2 (general kernel)
cudaError_t launch_tile_class(int itype, const void* rowptr, const TileDesc* tiles, i64 ntiles, int window, int cap, int rp_cap, int balance_pct,
                              unsigned char* cls, cudaStream_t st);
// hist[min(len, 1023)] += 1 for every row (hist: 1024 device counters, pre-zeroed)
cudaError_t launch_row_len_hist(int itype, const void* rowptr, i64 nrows, unsigned long long* hist, cudaStream_t st);
// rows longer than threshold: writes their local row ids (ascending not guaranteed) into rows_out (capacity cap), count via *count_out (device)
cudaError_t launch_find_long_rows(int itype, const void* rowptr, i64 nrows, i64 threshold, i64* rows_out, i64 cap,
                                  unsigned long long* count_out, cudaStream_t st);
// flags[t] = 1 iff tile t references a column outside [own_lo, own_lo+own_n)
cudaError_t launch_classify_tiles(int itype, const void* colval, const TileDesc* tiles, i64 ntiles, i64 own_lo, i64 own_n,
                                  unsigned char* flags, cudaStream_t st);
// out[t] = largest own column of tile t, 0-based relative to own_lo, or -1
cudaError_t launch_tile_maxcol(int itype, const void* colval, const TileDesc* tiles, i64 ntiles, i64 own_lo, i64 own_n, i64* out, cudaStream_t st);
cudaError_t launch_spmv_rowwalk(const SpmvLaunch& L, cudaStream_t st);  // tiles of class 1
cudaError_t launch_spmv_general(const SpmvLaunch& L, cudaStream_t st);  // tiles of class 2
// out2[0] = sum of the n per-CTA partials of a fused dot, in a fixed order; out2[1] = 0
cudaError_t launch_dot_partials_sum(const double* partials, i64 n, double* out2, cudaStream_t st);
cudaError_t launch_spmv_direct(const SpmvLaunch& L, cudaStream_t st);   // tiles of class 1, ghost-free, as runs of consecutive tiles
cudaError_t launch_build_tile_headers(int itype, const void* rowptr, const TileDesc* tiles, const unsigned char* cls, i64 ntiles, int window, int hdr_bytes,
                                      unsigned char* hdrs, cudaStream_t st);

// Sparse x dense: kn (1 or 4) columns k0 .. k0+kn of C = A * B over the tiles of `recs`.
struct SpmmLaunch {
    int dtype, itype;
    const void* rowptr;
    const void* colval;
    const void* nzval;
    i64 nrows, nnz;
    TileShape shape;
    const TileRec* recs;
    int n_runs = 0;
    int run_cta0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int run_tile0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_launch;
    const void* b_own;  // B's local block, column-major, at the first own source row (column 0)
    i64 ldb;
    const void* ghost;  // compact row-major ghost rows: ghost[g * ncols + k]
    int ncols;
    i64 own_lo, own_n;
    bool has_ghost;
    void* c;  // C's local block, column-major
    i64 ldc;
    int k0, kn;
};
bool spmm_supports_rowwalk(const TileShape& shape);
cudaError_t launch_spmm_rowwalk(const SpmmLaunch& L, cudaStream_t st);  // row-walk tiles
cudaError_t launch_spmm_rows(const SpmmLaunch& L, cudaStream_t st);     // any tiles, warp per row (general tiles, long rows)
// out[i * ncols + k] = B[idx[i] - 1 + k * ldb]
cudaError_t launch_pack_rows(int dtype, const void* B, i64 ldb, const i64* idx, i64 n, int ncols, void* out, cudaStream_t st);

// Row walk on compact tiles (compact.cu): per (matrix, plan) derived data of the interior row-walk tiles
struct CompactShape {
    int crows;       // rows a compact tile header can describe
    int chdr_bytes;  // bytes per tile header (multiple of 16)
    int chdr_fetch;  // bytes of it the multiply fetches
    int cw;          // staged entries per tile (window + over-fetch, multiple of 8)
    int cp_bytes;    // bytes of 16-bit positions per tile
    int xcap;        // staged x elements per tile
};
CompactShape compact_shape(int dtype, const TileShape& shape);
size_t cwalk_smem_bytes(int dtype, const CompactShape& sh);
// one CTA per listed tile: build == false: stats[i] = {x runs or -1, staged x elements}; build == true: headers + positions
cudaError_t launch_compact_tiles(bool build, int dtype, int itype, const void* rowptr, const void* colval, const TileDesc* tiles, const int* d_tile_ids, int n,
                                 int window, i64 own_lo, i64 own_n, const CompactShape& sh, int2* d_stats, unsigned char* d_hdrs, unsigned char* d_colpos,
                                 int* d_tail_q_min, cudaStream_t st);
struct CWalkLaunch {
    int dtype, lanes, window;
    const void* nzval;
    i64 nnz;
    CompactShape sh;
    const unsigned char* hdrs;
    const unsigned char* colpos;
    int q0;  // position of the first CTA's tile in the compact list
    int tail_q_min = 0x7fffffff;  // first position whose tile has tail elements of x
    int n_runs = 0;
    int run_cta0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int run_tile0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_launch;
    const void* x_own;  // first own element of x.v; 16-byte aligned
    void* y;
    const void* dot_x = nullptr;
    double* dot_out = nullptr;
};
cudaError_t launch_spmv_cwalk(const CWalkLaunch& L, cudaStream_t st);
// sparse x dense on the same compact tiles: kn (8, 4 or 1) columns from column k0
struct CWalkMLaunch {
    int dtype, lanes, window;
    const void* nzval;
    i64 nnz;
    CompactShape sh;
    const unsigned char* hdrs;
    const unsigned char* colpos;
    int q0;
    int tail_q_min = 0x7fffffff;
    int n_runs = 0;
    int run_cta0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int run_tile0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_launch;
    const void* b_own;  // B's local block at its first own row, column 0
    i64 ldb;
    void* c;  // C's local block, column 0
    i64 ldc;
    int k0, kn;
};
bool spmm_cwalk_supported(const CWalkMLaunch& L);
cudaError_t launch_spmm_cwalk(const CWalkMLaunch& L, cudaStream_t st);
// the same tiles behind a ring: persistent CTAs, producer warp + consumer warps (kn = 1: plain multiply, c = y, b_own = x.v)
cudaError_t launch_cring(const CWalkMLaunch& L, cudaStream_t st);

// nnz-split multiply for irregular matrices (flat.cu): per-matrix derived structure
struct FlatData {
    i64 n_chunks = 0;     // CTA chunks of 4096 stored entries (0: the matrix does not use the flat kernel)
    i64 n_wchunks = 0;    // warp chunks of 512 stored entries that hold entries
    i64 n_nonempty = 0, n_empty = 0;
    unsigned* d_bits = nullptr;   // row-start flags, one bit per stored entry (padded to whole chunks)
    i64* d_wrow = nullptr;        // per warp chunk: ordinal of the row holding its first entry
    i64* d_row_map = nullptr;     // ordinal -> row (only when some rows are empty)
    i64* d_empty_rows = nullptr;  // rows without entries (their y is zeroed)
    void* d_heads = nullptr;      // T[n_wchunks]: partial sums in front of each warp chunk's first row start
};
struct FlatLaunch {
    int dtype, itype;
    const void* colval;
    const void* nzval;
    i64 nnz;
    const FlatData* flat;
    const void* x_own;
    const void* gathered;
    i64 own_lo, own_n;
    bool has_ghost;
    bool keep_x;  // gather x with an L2 evict-last hint
    void* y;
    i64 long_threshold;
};
cudaError_t flat_build(int itype, const void* rowptr, i64 nrows, i64 nnz, FlatData* out, cudaStream_t st);
void flat_free(FlatData* F);
cudaError_t launch_spmv_flat(const FlatLaunch& L, cudaStream_t st);
// counts violations of: rowptr[0] == 1, rowptr non-decreasing, rowptr[nrows] == nnz + 1, 1 <= colval <= ncc
cudaError_t launch_validate_csr(int itype, const void* rowptr, const void* colval, i64 nrows, i64 nnz, i64 ncc, unsigned* d_bad, cudaStream_t st);

struct LongRowsLaunch {
    int dtype, itype;
    const void* rowptr;
    const void* colval;
    const void* nzval;
    const i64* long_rows;   // [nlong] local row ids
    const i64* chunk_ptr;   // [nlong+1] prefix of chunk counts
    i64 nlong, nchunks, chunk_nnz;
    const void* x_own;
    const void* gathered;
    i64 own_lo, own_n;
    bool has_ghost;
    void* partials;  // T[nchunks]
    void* y;
};
cudaError_t launch_long_rows(const LongRowsLaunch& L, cudaStream_t st);

// context.cu, for the other translation units: who am I, and an all-to-all of byte ranges between device buffers
// (grouped ncclSend/ncclRecv; the range for my own rank is a device-to-device copy).  NCCL world or nranks == 1.
int ctx_rank_info(const hpcla_ctx* ctx, int* device, int* rank, int* nranks, int* has_comm);
int ctx_agree_max(hpcla_ctx* ctx, int local, int* out, cudaStream_t stream);
int ctx_exchange_bytes(hpcla_ctx* ctx, const void* d_send, const i64* send_off, const i64* send_bytes, void* d_recv, const i64* recv_off,
                       const i64* recv_bytes, cudaStream_t stream);

// *flag = value after everything enqueued before it on the stream (system scope; the flag may live on a peer GPU)
cudaError_t launch_write_flag(unsigned* flag, unsigned value, cudaStream_t st);
cudaError_t preload_halo_kernels();
// out[k] = x[idx[k]-1]          (pack of src/vectors.jl:431-437, all peers in one launch)
cudaError_t launch_pack(int dtype, const void* x, const i64* idx, i64 n, void* out, cudaStream_t st);
// gathered[dst[k]-1] = x[src[k]-1]   (local copy of src/vectors.jl:426-428 == _gather_kernel! :174-177)
cudaError_t launch_local_copy(int dtype, const void* x, const i64* src, const i64* dst, i64 n, void* gathered, cudaStream_t st);

// reductions / updates.  scratch: device buffer of >= reduce_scratch_elems() doubles.
int reduce_scratch_doubles();
// out2[0..1] (double, device) = sum conj(x)*y (re, im); F32 accumulates in float per thread, double across threads
cudaError_t launch_dot(int dtype, i64 n, const void* x, const void* y, double* scratch, double* out2, cudaStream_t st);
cudaError_t launch_axpby(int dtype, i64 n, const void* alpha_host, const void* x, const void* beta_host, void* y, cudaStream_t st);

// CG building blocks (real types).  All scalars are (re, im) double pairs on the device.
cudaError_t launch_cg_init(int dtype, i64 n, const void* b, void* x, void* r, void* p, double* scratch, double* rr_out2, cudaStream_t st);
cudaError_t launch_cg_update_xr(int dtype, i64 n, const void* p, const void* q, void* x, void* r, const double* rr, const double* pq,
                                double* scratch, double* rr_new_out2, cudaStream_t st);
cudaError_t launch_cg_update_p(int dtype, i64 n, const void* r, void* p, const double* rr_new, const double* rr, cudaStream_t st);

}  // namespace hpcla
