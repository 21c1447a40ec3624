// compact.cu — the row walk on COMPACT TILES: the multiply for the interior tiles of stencil-like (banded) matrices.
//
// What the plain row walk (kernels.cu) leaves on the table, measured (profiles/r1j_ncu_full_poisson_rowwalk.txt): its
// lanes spend their time waiting on x gathers that go through L1TEX one 8-byte element at a time (long-scoreboard
// stalls 18.5 per issue), while the CTA's shared memory holds no copy in flight; and every stored entry moves a 4- or
// 8-byte column index although, inside one tile, the columns touched form a handful of short contiguous runs of x.
// Once per (matrix, plan) every interior row-walk tile therefore gets, next to the matrix (library-owned, like the
// tile headers of the direct walk):
//   * its X RUNS: <= 16 contiguous ranges of x.v that cover every column the tile reads (7-point stencil: 3 runs,
//     27-point: 9), found by sorting the tile's columns;
//   * its column indices re-expressed as 16-BIT POSITIONS into the staged runs (2 bytes per entry instead of 4 or 8),
//     and its row offsets as 16-bit values relative to the tile window.
// The multiply then fetches values (in place from A.nzval: in-place value writes are still seen), positions, header and
// the x runs with bulk asynchronous copies (1-D TMA) and walks the rows entirely out of shared memory: no gather leaves
// the SM.  Sums run left to right over separately rounded products with one lane per row, so y stays bit-identical to
// the reference's loop (src/sparse.jl:2055-2066).
// Bytes per stored entry of the 7-point Float64 case: 8 (value) + 2 (position) instead of 8 + 4.
#include <type_traits>

#include "device_common.cuh"

namespace hpcla {

constexpr int CW_R = 16;             // x runs per tile
constexpr int CW_HDR_FIXED = 32 + CW_R * 16;  // fixed part of a tile header: 32-byte head + the run table

struct CHead {  // 32 bytes
    i64 r0;
    int nrows, n_st;   // rows of the tile; staged entries (from the window start)
    int nruns, x_total;
    int tail_n, tail_soff;  // <= 3 elements at the very end of x.v that a 16-byte copy cannot fetch: count, position in sx
};
struct CRun {  // 16 bytes
    int xoff;  // first element of the run, relative to the own segment of x.v (multiple of 16 bytes)
    int len;   // elements fetched by the bulk copy (multiple of 16 bytes)
    int soff;  // position of the run in the staged x
    int tail_xoff;  // (run 0 only) x offset of the tail elements
};

template <class T>
struct CWalkArgs {
    const T* nzval;
    const unsigned char* hdrs;    // [n] tile headers, chdr_bytes apart, by position in the compact list
    const unsigned char* colpos;  // [n] 16-bit positions, cp_bytes apart
    const T* x_own;               // first own element of x.v (16-byte aligned)
    T* y;
    i64 nnz_total;
    TileRuns runs;
    int q0;  // position of this launch's first CTA in the compact list
    int tail_q_min;  // first position in the compact list whose tile has tail elements of x (INT_MAX: none)
    int window, cw, chdr_bytes, chdr_fetch, cp_bytes, xcap;
    const T* dot_x;
    double* dot_out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class T, int G, bool DOT>
__global__ void __launch_bounds__(ROW_THREADS, RowCfg<T>::CTAS) spmv_cwalk_kernel(const CWalkArgs<T> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem_raw);  // header + positions
    uint64_t* barB = barA + 1;                               // values
    uint64_t* barX = barA + 2;                               // x runs
    const unsigned char* shdr = smem_raw + 32;
    const unsigned short* spos = reinterpret_cast<const unsigned short*>(smem_raw + 32 + a.chdr_bytes);
    T* sval = reinterpret_cast<T*>(smem_raw + 32 + a.chdr_bytes + a.cp_bytes);
    T* sx = sval + a.cw;
    const int tid = threadIdx.x;
    const i64 tile = tile_of_cta(a.runs);
    const i64 q = (i64)a.q0 + blockIdx.x;
    const i64 w0 = tile * (i64)a.window;
    const i64 left = (a.nnz_total - w0) & ~(i64)3;
    const int n_fetch = (int)(left < (i64)a.cw ? (left > 0 ? left : 0) : (i64)a.cw);
    const unsigned char* ghdr = a.hdrs + q * (i64)a.chdr_bytes;
    if (tid == 0) {
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        mbar_init(barX, 32);
        mbar_fence_init();
        const uint64_t pol = l2_evict_first_policy();
        mbar_expect_tx(barA, (uint32_t)a.chdr_fetch + (uint32_t)a.cp_bytes);
        bulk_g2s(const_cast<unsigned char*>(shdr), ghdr, (uint32_t)a.chdr_fetch, barA, pol);
        bulk_g2s(const_cast<unsigned short*>(spos), a.colpos + q * (i64)a.cp_bytes, (uint32_t)a.cp_bytes, barA, pol);
        mbar_expect_tx(barB, (uint32_t)n_fetch * (uint32_t)sizeof(T));
        if (n_fetch > 0) bulk_g2s(sval, a.nzval + w0, (uint32_t)n_fetch * (uint32_t)sizeof(T), barB, pol);
    }
    __syncthreads();  // barriers initialised before anyone arrives or waits (nothing slow in front of this: no global load)
    if (tid < 32) {
        // the x runs: lane j reads run j of the header straight from global memory (in parallel with the copies above)
        // and issues its bulk copy; x is re-read by the neighbouring tiles, so it keeps the default L2 policy
        // (unused table entries are zero: no second load that would depend on the run count)
        int4 run = make_int4(0, 0, 0, 0);
        if (tid < CW_R) run = __ldg(reinterpret_cast<const int4*>(ghdr + 32) + tid);
        if (run.y > 0) {
            const uint32_t bytes = (uint32_t)run.y * (uint32_t)sizeof(T);
            mbar_expect_tx(barX, bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sx + run.z)),
                         "l"(a.x_own + run.x), "r"(bytes), "r"(smem_u32(barX))
                         : "memory");
        } else {
            mbar_arrive(barX);
        }
    }
    mbar_wait(barA, 0);
    const CHead* h = reinterpret_cast<const CHead*>(shdr);
    const i64 r0 = h->r0;
    const int nrows = h->nrows;
    const unsigned short* off = reinterpret_cast<const unsigned short*>(shdr + CW_HDR_FIXED);
    {
        // Rare and uniform per CTA: elements that no 16-byte copy may fetch — the <= 3 elements at the very end of x.v
        // (only tiles from tail_q_min on can have them) and the <= 3 values at the end of A.nzval (last tile).
        const i64 avail = a.nnz_total - w0;
        const int n_avail = (int)(avail < (i64)a.cw ? avail : (i64)a.cw);
        const bool x_tail = q >= (i64)a.tail_q_min, v_tail = n_fetch < n_avail;
        if (x_tail || v_tail) {
            if (x_tail && tid < h->tail_n) sx[h->tail_soff + tid] = a.x_own[reinterpret_cast<const CRun*>(shdr + 32)->tail_xoff + tid];
            if (v_tail)
                for (int k = n_fetch + tid; k < n_avail; k += ROW_THREADS) sval[k] = a.nzval[w0 + k];
            __syncthreads();
        }
    }
    constexpr int RPP = ROW_THREADS / G;
    constexpr int B = HPCLA_WALK_BATCH;
    const int lane = tid % G;
    double dot = 0.0;
    // fused dot (CG's p.q): this thread's element of p for its first row is requested now, so that the load flies while the
    // x runs and the values are still landing (it is the only global load of the walk)
    T dx0 = el_zero(T());
    if (DOT && lane == 0 && tid / G < nrows) dx0 = ld_x(a.dot_x + r0 + tid / G);
    mbar_wait(barX, 0);
    mbar_wait(barB, 0);
    for (int base = 0; base < nrows; base += RPP) {
        const int i = base + tid / G;
        const bool valid = i < nrows;
        const int b = valid ? (int)off[i] : 0;
        const int e = valid ? (int)off[i + 1] : 0;
        T dx = dx0;
        if (DOT && base > 0 && valid && lane == 0) dx = ld_x(a.dot_x + r0 + i);
        T acc = el_zero(T());
        for (int k = b + lane; k < e; k += B * G) {
            T p[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int kk = k + u * G;
                p[u] = (kk < e) ? el_mul(sval[kk], sx[spos[kk]]) : el_zero(T());
            }
#pragma unroll
            for (int u = 0; u < B; ++u)
                if (k + u * G < e) acc = el_add(acc, p[u]);
        }
        if (G > 1) {
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = el_add(acc, shfl_xor(acc, m));
        }
        if (valid && lane == 0) {
            st_y(a.y + r0 + i, acc);
            if (DOT) dot += dot_term(dx, acc);
        }
    }
    if (DOT) {  // CG's p.q rides on the multiply; one partial per CTA, summed in a fixed order later
        __shared__ double dsh[ROW_THREADS / 32];
        double v = dot;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if ((threadIdx.x & 31) == 0) dsh[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < ROW_THREADS / 32; ++w) t += dsh[w];
            a.dot_out[blockIdx.x] = t;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Sparse x dense on compact tiles: C[:, k0 .. k0+K) = A * B[:, k0 .. k0+K) for the interior tiles (A * B::HPCMatrix,
// src/sparse.jl:2391-2413).  B is column-major, so the x runs of a tile are contiguous in EVERY column: the same run table
// drives K bulk copies per run, the tile's values and 16-bit positions are fetched once for K right-hand sides, and the
// walk reads all K operands of an entry from shared memory (the plain kernel issues K separate 8-byte gathers per entry).
// ------------------------------------------------------------------------------------------------------------------
template <class T>
struct CWalkMArgs {
    const T* nzval;
    const unsigned char* hdrs;
    const unsigned char* colpos;
    const T* b_own;  // B's local block at its first own row, column k0 (16-byte aligned; ldb * sizeof(T) a multiple of 16)
    i64 ldb;
    T* c;  // C's local block, column k0
    i64 ldc;
    i64 nnz_total;
    TileRuns runs;
    int q0, tail_q_min;
    int window, cw, chdr_bytes, chdr_fetch, cp_bytes, xcap;
};

__device__ __forceinline__ void mbar_add_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <class T, int G, int K>
__global__ void __launch_bounds__(ROW_THREADS, K >= 8 ? 2 : K >= 4 ? 3 : RowCfg<T>::CTAS) spmm_cwalk_kernel(const CWalkMArgs<T> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* barB = barA + 1;
    uint64_t* barX = barA + 2;
    const unsigned char* shdr = smem_raw + 32;
    const unsigned short* spos = reinterpret_cast<const unsigned short*>(smem_raw + 32 + a.chdr_bytes);
    T* sval = reinterpret_cast<T*>(smem_raw + 32 + a.chdr_bytes + a.cp_bytes);
    T* sx = sval + a.cw;  // [K][xcap]
    const int tid = threadIdx.x;
    const i64 tile = tile_of_cta(a.runs);
    const i64 q = (i64)a.q0 + blockIdx.x;
    const i64 w0 = tile * (i64)a.window;
    const i64 left = (a.nnz_total - w0) & ~(i64)3;
    const int n_fetch = (int)(left < (i64)a.cw ? (left > 0 ? left : 0) : (i64)a.cw);
    const unsigned char* ghdr = a.hdrs + q * (i64)a.chdr_bytes;
    if (tid == 0) {
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        mbar_init(barX, 32);
        mbar_fence_init();
        const uint64_t pol = l2_evict_first_policy();
        mbar_expect_tx(barA, (uint32_t)a.chdr_fetch + (uint32_t)a.cp_bytes);
        bulk_g2s(const_cast<unsigned char*>(shdr), ghdr, (uint32_t)a.chdr_fetch, barA, pol);
        bulk_g2s(const_cast<unsigned short*>(spos), a.colpos + q * (i64)a.cp_bytes, (uint32_t)a.cp_bytes, barA, pol);
        mbar_expect_tx(barB, (uint32_t)n_fetch * (uint32_t)sizeof(T));
        if (n_fetch > 0) bulk_g2s(sval, a.nzval + w0, (uint32_t)n_fetch * (uint32_t)sizeof(T), barB, pol);
    }
    __syncthreads();
    if (tid < 32) {  // K copies per x run, spread over the lanes of warp 0
#pragma unroll
        for (int idx = tid; idx < K * CW_R; idx += 32) {
            const int k = idx / CW_R, j = idx % CW_R;
            const int4 run = __ldg(reinterpret_cast<const int4*>(ghdr + 32) + j);
            if (run.y > 0) {
                const uint32_t bytes = (uint32_t)run.y * (uint32_t)sizeof(T);
                mbar_add_tx(barX, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(sx + (size_t)k * a.xcap + run.z)),
                             "l"(a.b_own + (i64)k * a.ldb + run.x), "r"(bytes), "r"(smem_u32(barX))
                             : "memory");
            }
        }
        mbar_arrive(barX);
    }
    mbar_wait(barA, 0);
    const CHead* h = reinterpret_cast<const CHead*>(shdr);
    const i64 r0 = h->r0;
    const int nrows = h->nrows;
    const unsigned short* off = reinterpret_cast<const unsigned short*>(shdr + CW_HDR_FIXED);
    {
        const i64 avail = a.nnz_total - w0;
        const int n_avail = (int)(avail < (i64)a.cw ? avail : (i64)a.cw);
        const bool x_tail = q >= (i64)a.tail_q_min, v_tail = n_fetch < n_avail;
        if (x_tail || v_tail) {
            if (x_tail && tid < h->tail_n * K) {
                const int k = tid / h->tail_n, t = tid % h->tail_n;
                sx[(size_t)k * a.xcap + h->tail_soff + t] = a.b_own[(i64)k * a.ldb + reinterpret_cast<const CRun*>(shdr + 32)->tail_xoff + t];
            }
            if (v_tail)
                for (int kk = n_fetch + tid; kk < n_avail; kk += ROW_THREADS) sval[kk] = a.nzval[w0 + kk];
            __syncthreads();
        }
    }
    mbar_wait(barX, 0);
    mbar_wait(barB, 0);
    constexpr int RPP = ROW_THREADS / G;
    const int lane = tid % G;
    for (int base = 0; base < nrows; base += RPP) {
        const int i = base + tid / G;
        const bool valid = i < nrows;
        const int b = valid ? (int)off[i] : 0;
        const int e = valid ? (int)off[i + 1] : 0;
        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = el_zero(T());
        for (int kk = b + lane; kk < e; kk += G) {
            const T v = sval[kk];
            const T* xp = sx + spos[kk];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], el_mul(v, xp[(size_t)k * a.xcap]));
        }
        if (G > 1) {
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1)
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], shfl_xor(acc[k], m));
        }
        if (valid && lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) st_y(a.c + r0 + i + (i64)k * a.ldc, acc[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// The same compact tiles behind a RING: one persistent CTA per SM, a producer warp that keeps S tiles in flight (bulk
// copies of header, positions, values and the x runs of K columns into stage it mod S, completing on full[s]) and eight
// consumer warps that walk tile after tile out of shared memory and hand each stage back through empty[s].  The copies
// of the next tiles proceed while a tile is walked, whatever the occupancy — which is what the one-tile-per-CTA kernels
// above lose when a tile needs a large share of the SM's shared memory (sparse x dense: 100 KB per tile at 8 columns,
// two CTAs per SM, nothing in flight while both walk).  K = 1 is the plain multiply.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RING_THREADS = ROW_THREADS + 32;

__device__ __forceinline__ i64 tile_of_pos(const TileRuns& runs, int pos) {
    int skip = runs.skip[0];
    if (runs.n > 1) {
#pragma unroll
        for (int j = 1; j < 8; ++j) skip = (pos >= runs.cta0[j]) ? runs.skip[j] : skip;
    }
    return (i64)(pos + skip);
}

template <class T, int G, int K>
__global__ void __launch_bounds__(RING_THREADS, 1) cring_kernel(const CWalkMArgs<T> a, int n_tiles, int stages, int stage_bytes) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [stages]
    uint64_t* empty = full + 8;                              // [stages]  (stages <= 8)
    unsigned char* ring = smem_raw + 128;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full + s, 32);                // the producer's lanes
            mbar_init(empty + s, ROW_THREADS / 32);  // one arrival per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int grid = (int)gridDim.x;
    if (tid >= ROW_THREADS) {
        // ---------------- producer warp ----------------
        const int lane = tid - ROW_THREADS;
        const uint64_t pol = l2_evict_first_policy();
        int pos = (int)blockIdx.x;
        int4 run = make_int4(0, 0, 0, 0), run_next = make_int4(0, 0, 0, 0);
        int4 head_tail = make_int4(0, 0, 0, 0);
        if (pos < n_tiles) {
            const unsigned char* gh = a.hdrs + ((i64)a.q0 + pos) * (i64)a.chdr_bytes;
            run = __ldg(reinterpret_cast<const int4*>(gh + 32) + (lane & (CW_R - 1)));
        }
        for (int it = 0; pos < n_tiles; ++it, pos += grid) {
            const int s = it % stages;
            const i64 q = (i64)a.q0 + pos;
            const unsigned char* ghdr = a.hdrs + q * (i64)a.chdr_bytes;
            if (pos + grid < n_tiles)  // the next tile's run table is requested now: its latency hides behind this tile's issue
                run_next = __ldg(reinterpret_cast<const int4*>(a.hdrs + (q + grid) * (i64)a.chdr_bytes + 32) + (lane & (CW_R - 1)));
            const bool x_tail = q >= (i64)a.tail_q_min;
            if (x_tail) head_tail = __ldg(reinterpret_cast<const int4*>(ghdr) + 1);  // {nruns, x_total, tail_n, tail_soff}
            if (it >= stages) mbar_wait(empty + s, (uint32_t)((it / stages - 1) & 1));
            unsigned char* st = ring + (size_t)s * stage_bytes;
            unsigned short* spos = reinterpret_cast<unsigned short*>(st + a.chdr_bytes);
            T* sval = reinterpret_cast<T*>(st + a.chdr_bytes + a.cp_bytes);
            T* sx = sval + a.cw;
            const i64 tile = tile_of_pos(a.runs, pos);
            const i64 w0 = tile * (i64)a.window;
            const i64 left = (a.nnz_total - w0) & ~(i64)3;
            const int n_fetch = (int)(left < (i64)a.cw ? (left > 0 ? left : 0) : (i64)a.cw);
            if (lane == 0) {
                mbar_add_tx(full + s, (uint32_t)a.chdr_fetch + (uint32_t)a.cp_bytes + (uint32_t)n_fetch * (uint32_t)sizeof(T));
                bulk_g2s(st, ghdr, (uint32_t)a.chdr_fetch, full + s, pol);
                bulk_g2s(spos, a.colpos + q * (i64)a.cp_bytes, (uint32_t)a.cp_bytes, full + s, pol);
                if (n_fetch > 0) bulk_g2s(sval, a.nzval + w0, (uint32_t)n_fetch * (uint32_t)sizeof(T), full + s, pol);
            }
            if (run.y > 0) {
                const uint32_t bytes = (uint32_t)run.y * (uint32_t)sizeof(T);
#pragma unroll
                for (int idx = lane; idx < K * CW_R; idx += 32) {
                    const int k = idx / CW_R;
                    mbar_add_tx(full + s, bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_u32(sx + (size_t)k * a.xcap + run.z)),
                                 "l"(a.b_own + (i64)k * a.ldb + run.x), "r"(bytes), "r"(smem_u32(full + s))
                                 : "memory");
                }
            }
            {  // rare: elements no 16-byte copy may fetch (end of x.v: tiles from tail_q_min on; end of A.nzval: last tile)
                const i64 avail = a.nnz_total - w0;
                const int n_avail = (int)(avail < (i64)a.cw ? avail : (i64)a.cw);
                if (x_tail && head_tail.z > 0) {
                    const int tail_xoff = __ldg(reinterpret_cast<const int*>(ghdr + 32) + 3);
                    for (int e = lane; e < head_tail.z * K; e += 32) {
                        const int k = e / head_tail.z, t = e % head_tail.z;
                        sx[(size_t)k * a.xcap + head_tail.w + t] = a.b_own[(i64)k * a.ldb + tail_xoff + t];
                    }
                }
                for (int kk = n_fetch + lane; kk < n_avail; kk += 32) sval[kk] = a.nzval[w0 + kk];
            }
            mbar_arrive(full + s);  // (release: the scalar stores above are visible to whoever sees the phase complete)
            run = run_next;
        }
    } else {
        // ---------------- consumer warps ----------------
        constexpr int RPP = ROW_THREADS / G;
        const int lane = tid % G;
        int it = 0;
        for (int pos = (int)blockIdx.x; pos < n_tiles; ++it, pos += grid) {
            const int s = it % stages;
            const unsigned char* st = ring + (size_t)s * stage_bytes;
            const unsigned short* spos = reinterpret_cast<const unsigned short*>(st + a.chdr_bytes);
            const T* sval = reinterpret_cast<const T*>(st + a.chdr_bytes + a.cp_bytes);
            const T* sx = sval + a.cw;
            mbar_wait(full + s, (uint32_t)((it / stages) & 1));
            const CHead* h = reinterpret_cast<const CHead*>(st);
            const i64 r0 = h->r0;
            const int nrows = h->nrows;
            const unsigned short* off = reinterpret_cast<const unsigned short*>(st + CW_HDR_FIXED);
            for (int base = 0; base < nrows; base += RPP) {
                const int i = base + tid / G;
                const bool valid = i < nrows;
                const int b = valid ? (int)off[i] : 0;
                const int e = valid ? (int)off[i + 1] : 0;
                T acc[K];
#pragma unroll
                for (int k = 0; k < K; ++k) acc[k] = el_zero(T());
                if (K == 1) {
                    constexpr int B = HPCLA_WALK_BATCH;
                    for (int kk = b + lane; kk < e; kk += B * G) {
                        T p[B];
#pragma unroll
                        for (int u = 0; u < B; ++u) {
                            const int k2 = kk + u * G;
                            p[u] = (k2 < e) ? el_mul(sval[k2], sx[spos[k2]]) : el_zero(T());
                        }
#pragma unroll
                        for (int u = 0; u < B; ++u)
                            if (kk + u * G < e) acc[0] = el_add(acc[0], p[u]);
                    }
                } else {
                    for (int kk = b + lane; kk < e; kk += G) {
                        const T v = sval[kk];
                        const T* xp = sx + spos[kk];
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], el_mul(v, xp[(size_t)k * a.xcap]));
                    }
                }
                if (G > 1) {
#pragma unroll
                    for (int m = G / 2; m >= 1; m >>= 1)
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], shfl_xor(acc[k], m));
                }
                if (valid && lane == 0) {
#pragma unroll
                    for (int k = 0; k < K; ++k) st_y(a.c + r0 + i + (i64)k * a.ldc, acc[k]);
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(empty + s);  // this warp is done with the stage
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// set-up: analysis (BUILD = false: stats[i] = {runs, staged x elements}, runs = -1 when the tile cannot be compacted) and
// construction (BUILD = true) of the compact data of one tile per CTA.  The tile's columns are sorted in shared memory
// (bitonic), cut into runs wherever two neighbours are more than 64 bytes apart, and every entry's column is replaced by
// its position in the staged runs.
// ------------------------------------------------------------------------------------------------------------------
template <class Ti, bool BUILD>
__global__ void __launch_bounds__(256) compact_tile_kernel(const Ti* __restrict__ rowptr, const Ti* __restrict__ colval, const TileDesc* __restrict__ tiles,
                                                           const int* __restrict__ tile_ids, int window, i64 own_lo, i64 own_n, int elem_bytes,
                                                           CompactShape sh, int p2, int2* __restrict__ stats, unsigned char* __restrict__ hdrs,
                                                           unsigned char* __restrict__ colpos, int* __restrict__ tail_q_min) {
    extern __shared__ unsigned int keys[];  // [p2]
    __shared__ int run_xoff[CW_R + 1], run_len[CW_R + 1], run_soff[CW_R + 1];
    __shared__ int s_nruns, s_total, s_tail_n, s_tail_soff, s_tail_xoff;
    const int tid = threadIdx.x;
    const i64 t = tile_ids[blockIdx.x];
    const i64 r0 = tiles[t].row, r1 = tiles[t + 1].row, e = tiles[t + 1].nnz, w0 = t * (i64)window;
    const int n_st = (int)(e - w0), nrows = (int)(r1 - r0);
    const int A = 16 / elem_bytes;  // elements per 16 bytes (1 for ComplexF64)
    const bool shape_ok = n_st <= sh.cw && n_st > 0 && nrows <= sh.crows && n_st <= p2;
    if (!shape_ok) {
        if (!BUILD && tid == 0) stats[blockIdx.x] = make_int2(-1, 0);
        return;
    }
    for (int k = tid; k < p2; k += 256) keys[k] = k < n_st ? (unsigned int)((i64)colval[w0 + k] - own_lo) : 0xffffffffu;
    __syncthreads();
    for (int size = 2; size <= p2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < p2 / 2; i += 256) {
                const int lo = 2 * i - (i & (stride - 1));  // index of the lower element of the pair
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned int x = keys[lo], y = keys[hi];
                if ((x > y) == up) keys[lo] = y, keys[hi] = x;
            }
            __syncthreads();
        }
    if (tid == 0) {
        const int gap_groups = 64 / 16;  // neighbours more than 64 bytes apart start a new run
        int nruns = 0, total = 0;
        unsigned int g_first = 0, g_prev = 0;
        bool ok = true;
        for (int i = 0; i <= n_st && ok; ++i) {
            const unsigned int g = i < n_st ? keys[i] / (unsigned)A : 0u;
            if (i == n_st || (i > 0 && g - g_prev > (unsigned)gap_groups) || i == 0) {
                if (i > 0) {  // close the run [g_first, g_prev]
                    if (nruns == CW_R) ok = false;
                    else {
                        run_xoff[nruns] = (int)(g_first * (unsigned)A);
                        run_len[nruns] = (int)((g_prev - g_first + 1) * (unsigned)A);
                        run_soff[nruns] = total;
                        total += run_len[nruns];
                        ++nruns;
                    }
                }
                g_first = g;
            }
            g_prev = g;
        }
        if (total > 65535 || (BUILD && total > sh.xcap)) ok = false;
        s_tail_n = 0, s_tail_soff = 0, s_tail_xoff = 0;
        if (ok && nruns > 0) {  // the last run may reach past the end of x.v by less than 16 bytes
            const int j = nruns - 1;
            if ((i64)run_xoff[j] + run_len[j] > own_n) {
                const int fetch = (int)(((own_n - run_xoff[j]) / A) * A);
                s_tail_n = (int)(own_n - run_xoff[j] - fetch);
                s_tail_soff = run_soff[j] + fetch;
                s_tail_xoff = run_xoff[j] + fetch;
                run_len[j] = fetch;  // (the position of the following data in sx is unchanged: the run is the last one)
            }
        }
        s_nruns = ok ? nruns : -1;
        s_total = total;
    }
    __syncthreads();
    if (!BUILD) {
        if (tid == 0) stats[blockIdx.x] = make_int2(s_nruns, s_total);
        return;
    }
    if (s_nruns < 0) return;  // (cannot happen for the tiles handed to the build pass)
    unsigned char* h = hdrs + (i64)blockIdx.x * sh.chdr_bytes;
    if (tid == 0) {
        CHead* hd = reinterpret_cast<CHead*>(h);
        hd->r0 = r0;
        hd->nrows = nrows;
        hd->n_st = n_st;
        hd->nruns = s_nruns;
        hd->x_total = s_total;
        hd->tail_n = s_tail_n;
        hd->tail_soff = s_tail_soff;
        if (s_tail_n > 0 && tail_q_min) atomicMin(tail_q_min, (int)blockIdx.x);
    }
    if (tid < CW_R) {
        CRun* rn = reinterpret_cast<CRun*>(h + 32) + tid;
        const bool in = tid < s_nruns;
        rn->xoff = in ? run_xoff[tid] : 0;
        rn->len = in ? run_len[tid] : 0;
        rn->soff = in ? run_soff[tid] : 0;
        rn->tail_xoff = tid == 0 ? s_tail_xoff : 0;
    }
    unsigned short* off = reinterpret_cast<unsigned short*>(h + CW_HDR_FIXED);
    for (int i = tid; i <= nrows; i += 256) off[i] = (unsigned short)((i64)rowptr[r0 + i] - 1 - w0);
    unsigned short* cp = reinterpret_cast<unsigned short*>(colpos + (i64)blockIdx.x * sh.cp_bytes);
    const int nruns = s_nruns;
    for (int k = tid; k < sh.cp_bytes / 2; k += 256) {
        unsigned short pos = 0;
        if (k < n_st) {
            const int c = (int)((i64)colval[w0 + k] - own_lo);
            int j = 0;
            while (j + 1 < nruns && c >= run_xoff[j + 1]) ++j;  // runs ascend
            pos = (unsigned short)(run_soff[j] + (c - run_xoff[j]));
        }
        cp[k] = pos;
    }
}

CompactShape compact_shape(int dtype, const TileShape& s) {
    CompactShape c{};
    if (s.lanes <= 0) return c;
    const int rpp = s.threads / s.lanes;
    c.crows = ((rpp + rpp / 4) + 7) & ~7;              // boundary rows of a stencil are shorter: up to 1.25x the typical row count
    if (c.crows > s.hdr_rows && s.hdr_rows > 0) c.crows = s.hdr_rows;
    c.chdr_fetch = (CW_HDR_FIXED + 2 * (c.crows + 1) + 15) & ~15;
    c.chdr_bytes = c.chdr_fetch;
    c.cw = (s.window + s.ovf + 7) & ~7;
    c.cp_bytes = 2 * c.cw;  // multiple of 16
    c.xcap = 0;             // chosen from the analysis pass
    (void)dtype;
    return c;
}

static int pow2_at_least(int n) {
    int p = 256;
    while (p < n) p <<= 1;
    return p;
}

cudaError_t launch_compact_tiles(bool build, int dtype, int itype, const void* rowptr, const void* colval, const TileDesc* tiles, const int* d_tile_ids, int n,
                                 int window, i64 own_lo, i64 own_n, const CompactShape& sh, int2* d_stats, unsigned char* d_hdrs, unsigned char* d_colpos,
                                 int* d_tail_q_min, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int p2 = pow2_at_least(sh.cw);
    if (p2 > 8192) return cudaErrorInvalidValue;
    const size_t smem = (size_t)p2 * 4;
    const int eb = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16;
#define CT_LAUNCH(Ti, B) compact_tile_kernel<Ti, B><<<n, 256, smem, st>>>((const Ti*)rowptr, (const Ti*)colval, tiles, d_tile_ids, window, own_lo, own_n, eb, sh, p2, d_stats, d_hdrs, d_colpos, d_tail_q_min)
    if (itype == HPCLA_I32) {
        if (build) CT_LAUNCH(int, true);
        else CT_LAUNCH(int, false);
    } else {
        if (build) CT_LAUNCH(long long, true);
        else CT_LAUNCH(long long, false);
    }
#undef CT_LAUNCH
    return cudaGetLastError();
}

size_t cwalk_smem_bytes(int dtype, const CompactShape& sh) {
    const size_t ts = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16;
    return 32 + (size_t)sh.chdr_bytes + (size_t)sh.cp_bytes + (size_t)(sh.cw + sh.xcap) * ts;
}

template <class T, int G>
static cudaError_t cwalk_launch(const CWalkLaunch& L, const CWalkArgs<T>& a, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if constexpr (!std::is_same<T, cplx>::value) {
        if (L.dot_x) {
            if ((e = ensure_smem<spmv_cwalk_kernel<T, G, true>>(smem, true)) != cudaSuccess) return e;
            spmv_cwalk_kernel<T, G, true><<<L.n_launch, ROW_THREADS, smem, st>>>(a);
            return cudaGetLastError();
        }
    }
    if ((e = ensure_smem<spmv_cwalk_kernel<T, G, false>>(smem, true)) != cudaSuccess) return e;
    spmv_cwalk_kernel<T, G, false><<<L.n_launch, ROW_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <class T>
static cudaError_t cwalk_typed(const CWalkLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    CWalkArgs<T> a;
    a.nzval = (const T*)L.nzval;
    a.hdrs = L.hdrs;
    a.colpos = L.colpos;
    a.x_own = (const T*)L.x_own;
    a.y = (T*)L.y;
    a.nnz_total = L.nnz;
    a.runs = launch_runs(L.n_runs, L.run_cta0, L.run_tile0);
    a.q0 = L.q0;
    a.tail_q_min = L.tail_q_min;
    a.window = L.window;
    a.cw = L.sh.cw;
    a.chdr_bytes = L.sh.chdr_bytes;
    a.chdr_fetch = L.sh.chdr_fetch;
    a.cp_bytes = L.sh.cp_bytes;
    a.xcap = L.sh.xcap;
    a.dot_x = (const T*)L.dot_x;
    a.dot_out = L.dot_out;
    const size_t smem = cwalk_smem_bytes(L.dtype, L.sh);
    switch (L.lanes) {
        case 1: return cwalk_launch<T, 1>(L, a, smem, st);
        case 2: return cwalk_launch<T, 2>(L, a, smem, st);
        case 4: return cwalk_launch<T, 4>(L, a, smem, st);
        case 8: return cwalk_launch<T, 8>(L, a, smem, st);
        case 16: return cwalk_launch<T, 16>(L, a, smem, st);
        case 32: return cwalk_launch<T, 32>(L, a, smem, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_spmv_cwalk(const CWalkLaunch& L, cudaStream_t st) {
    if (L.dtype == HPCLA_F32) return cwalk_typed<float>(L, st);
    if (L.dtype == HPCLA_F64) return cwalk_typed<double>(L, st);
    if (L.dtype == HPCLA_C128) return cwalk_typed<cplx>(L, st);
    return cudaErrorInvalidValue;
}

size_t cwalk_m_smem_bytes(int dtype, const CompactShape& sh, int K) {
    const size_t ts = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16;
    return 32 + (size_t)sh.chdr_bytes + (size_t)sh.cp_bytes + ((size_t)sh.cw + (size_t)K * sh.xcap) * ts;
}

template <class T, int G, int K>
static cudaError_t cwalk_m_launch(const CWalkMLaunch& L, const CWalkMArgs<T>& a, cudaStream_t st) {
    const size_t smem = cwalk_m_smem_bytes(L.dtype, L.sh, K);
    cudaError_t e;
    if ((e = ensure_smem<spmm_cwalk_kernel<T, G, K>>(smem, true)) != cudaSuccess) return e;
    spmm_cwalk_kernel<T, G, K><<<L.n_launch, ROW_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <class T, int K>
static cudaError_t cwalk_m_lanes(const CWalkMLaunch& L, cudaStream_t st) {
    const size_t es = sizeof(T);
    CWalkMArgs<T> a;
    a.nzval = (const T*)L.nzval;
    a.hdrs = L.hdrs;
    a.colpos = L.colpos;
    a.b_own = (const T*)L.b_own + (i64)L.k0 * L.ldb;
    a.ldb = L.ldb;
    a.c = (T*)L.c + (i64)L.k0 * L.ldc;
    a.ldc = L.ldc;
    a.nnz_total = L.nnz;
    a.runs = launch_runs(L.n_runs, L.run_cta0, L.run_tile0);
    a.q0 = L.q0;
    a.tail_q_min = L.tail_q_min;
    a.window = L.window;
    a.cw = L.sh.cw;
    a.chdr_bytes = L.sh.chdr_bytes;
    a.chdr_fetch = L.sh.chdr_fetch;
    a.cp_bytes = L.sh.cp_bytes;
    a.xcap = L.sh.xcap;
    (void)es;
    switch (L.lanes) {
        case 1: return cwalk_m_launch<T, 1, K>(L, a, st);
        case 2: return cwalk_m_launch<T, 2, K>(L, a, st);
        case 4: return cwalk_m_launch<T, 4, K>(L, a, st);
        case 8: return cwalk_m_launch<T, 8, K>(L, a, st);
    }
    return cudaErrorInvalidValue;
}

template <class T>
static cudaError_t cwalk_m_typed(const CWalkMLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    if (L.kn == 8) return cwalk_m_lanes<T, 8>(L, st);
    if (L.kn == 4) return cwalk_m_lanes<T, 4>(L, st);
    if (L.kn == 1) return cwalk_m_lanes<T, 1>(L, st);
    return cudaErrorInvalidValue;
}

// ring launch: stages = as many tiles as fit the SM's shared memory (at most 8)
template <class T, int G, int K>
static cudaError_t cring_launch(const CWalkMLaunch& L, const CWalkMArgs<T>& a, cudaStream_t st) {
    const size_t stage = cwalk_m_smem_bytes(L.dtype, L.sh, K) - 32;  // multiple of 16
    int stages = (int)((size_t)(226 * 1024 - 128) / stage);
    if (stages > 8) stages = 8;
    if (stages < 2) return cudaErrorInvalidValue;
    const size_t smem = 128 + (size_t)stages * stage;
    cudaError_t e;
    if ((e = ensure_smem<cring_kernel<T, G, K>>(smem, true)) != cudaSuccess) return e;
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    int grid = sms[dev & 63] > 0 ? sms[dev & 63] : 148;
    if (grid > L.n_launch) grid = L.n_launch;
    cring_kernel<T, G, K><<<grid, RING_THREADS, smem, st>>>(a, L.n_launch, stages, (int)stage);
    return cudaGetLastError();
}

template <class T, int K>
static cudaError_t cring_lanes(const CWalkMLaunch& L, cudaStream_t st) {
    CWalkMArgs<T> a;
    a.nzval = (const T*)L.nzval;
    a.hdrs = L.hdrs;
    a.colpos = L.colpos;
    a.b_own = (const T*)L.b_own + (i64)L.k0 * L.ldb;
    a.ldb = L.ldb;
    a.c = (T*)L.c + (i64)L.k0 * L.ldc;
    a.ldc = L.ldc;
    a.nnz_total = L.nnz;
    a.runs = launch_runs(L.n_runs, L.run_cta0, L.run_tile0);
    a.q0 = L.q0;
    a.tail_q_min = L.tail_q_min;
    a.window = L.window;
    a.cw = L.sh.cw;
    a.chdr_bytes = L.sh.chdr_bytes;
    a.chdr_fetch = L.sh.chdr_fetch;
    a.cp_bytes = L.sh.cp_bytes;
    a.xcap = L.sh.xcap;
    switch (L.lanes) {
        case 1: return cring_launch<T, 1, K>(L, a, st);
        case 2: return cring_launch<T, 2, K>(L, a, st);
        case 4: return cring_launch<T, 4, K>(L, a, st);
        case 8: return cring_launch<T, 8, K>(L, a, st);
    }
    return cudaErrorInvalidValue;
}

template <class T>
static cudaError_t cring_typed(const CWalkMLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    if (L.kn == 8) return cring_lanes<T, 8>(L, st);
    if (L.kn == 4) return cring_lanes<T, 4>(L, st);
    if (L.kn == 1) return cring_lanes<T, 1>(L, st);
    return cudaErrorInvalidValue;
}

cudaError_t launch_cring(const CWalkMLaunch& L, cudaStream_t st) {
    if (L.dtype == HPCLA_F32) return cring_typed<float>(L, st);
    if (L.dtype == HPCLA_F64) return cring_typed<double>(L, st);
    if (L.dtype == HPCLA_C128) return cring_typed<cplx>(L, st);
    return cudaErrorInvalidValue;
}

// whether the compact kernel can take this launch: staged x for kn columns must fit the SM's shared memory, the lanes must
// be a supported count, and every column's x runs must start on a 16-byte boundary
bool spmm_cwalk_supported(const CWalkMLaunch& L) {
    const size_t es = L.dtype == HPCLA_F32 ? 4 : L.dtype == HPCLA_F64 ? 8 : 16;
    if (L.lanes < 1 || L.lanes > 8) return false;
    if ((((uintptr_t)L.b_own) & 15) || ((L.ldb * (i64)es) & 15) || ((((uintptr_t)L.b_own) + (size_t)L.k0 * (size_t)L.ldb * es) & 15)) return false;
    return cwalk_m_smem_bytes(L.dtype, L.sh, L.kn) <= 110 * 1024;
}

cudaError_t launch_spmm_cwalk(const CWalkMLaunch& L, cudaStream_t st) {
    if (L.dtype == HPCLA_F32) return cwalk_m_typed<float>(L, st);
    if (L.dtype == HPCLA_F64) return cwalk_m_typed<double>(L, st);
    if (L.dtype == HPCLA_C128) return cwalk_m_typed<cplx>(L, st);
    return cudaErrorInvalidValue;
}

}  // namespace hpcla
