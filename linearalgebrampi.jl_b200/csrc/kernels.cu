// kernels.cu — hand-written sm_100a kernels of the distributed SpMV hot path.
//
// The arithmetic replaces the reference's one-work-item-per-row KernelAbstractions kernel
// (_spmv_kernel!, src/sparse.jl:2055-2066) and its gather kernel (_gather_kernel!, src/vectors.jl:174-177).
// SpMV is HBM-bound (0.10-0.37 flop/B): no tensor cores.  The local rows are cut into TILES (the rows whose first
// stored entry lies in one window of the nonzero stream); every tile is classified once per matrix:
//   * row-walk tiles (stencil-like: balanced rows) -> spmv_rowwalk_kernel: the tile's colval / nzval / rowptr slices
//     are staged by three bulk asynchronous copies (1-D TMA, evict-first in L2), then G lanes per row walk the staged
//     row and gather x through the read-only path; lanes of a warp own consecutive rows, so the gathers coalesce;
//   * general tiles -> spmv_tile_kernel: 128-bit streaming loads of consecutive NONZEROS per lane, products staged in
//     shared memory, per-row reduction by 1..32 lanes, warp-per-row for rows too long to stage;
//   * rows above the split threshold (power-law tails) -> chunk + ordered partial sums, no atomics.
// The reference's 1-based Int32/Int64 arrays are consumed as they are (no conversion pass, no private copy).
// With one lane per row the sum runs left to right over products rounded separately from the adds, i.e. it is
// bit-identical to the reference's `acc += nzval[j]*x[colval[j]]` (no FMA contraction).
#include <type_traits>

#include "device_common.cuh"

namespace hpcla {

// per-row reduction of staged products by G cooperating lanes (G = 1: the reference's left-to-right order)
template <class T, class Ti, int THREADS, int G>
__device__ __forceinline__ void reduce_rows(const T* prod, const Ti* __restrict__ rowptr, T* __restrict__ y, i64 r0, i64 r1, i64 s4,
                                            int tid) {
    constexpr int RPP = THREADS / G;  // rows per pass
    const int lane = tid % G;
    for (i64 base = r0; base < r1; base += RPP) {
        const i64 r = base + tid / G;
        const bool valid = r < r1;
        T acc = el_zero(T());
        if (valid) {
            const int b = (int)((i64)__ldg(rowptr + r) - 1 - s4);
            const int e = (int)((i64)__ldg(rowptr + r + 1) - 1 - s4);
            for (int k = b + lane; k < e; k += G) acc = el_add(acc, prod[k]);
        }
        if (G > 1) {
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = el_add(acc, shfl_xor(acc, m));
        }
        if (valid && lane == 0) st_y(y + r, acc);
    }
}

template <class T, class Ti, int THREADS, int GROUPS, bool GHOST>
__global__ void __launch_bounds__(THREADS, HPCLA_GENERAL_MIN_CTAS) spmv_tile_kernel(const TileArgs<T, Ti> a, int smem_elems) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* prod = reinterpret_cast<T*>(smem_raw);
    constexpr int CHUNK = THREADS * GROUPS * 4;
    const int tid = threadIdx.x;
    // one 32-byte record per CTA: no list -> descriptor -> data chain of dependent loads in front of the first copy
    const longlong2 d0 = __ldg(reinterpret_cast<const longlong2*>(a.recs + blockIdx.x));
    const longlong2 d1 = __ldg(reinterpret_cast<const longlong2*>(a.recs + blockIdx.x) + 1);
    const i64 r0 = d0.x, r1 = d0.y;
    if (r1 <= r0) return;
    const i64 s = d1.x, e = d1.y;  // 0-based nonzero range of the tile
    const i64 s4 = s & ~(i64)3;    // 16-byte aligned start for every element width
    const i64 n = e - s4;

    if (n <= (i64)smem_elems) {
        // ---- stage: coalesced 128-bit streaming loads of 4 consecutive nonzeros per lane, gather x, multiply ----
        const Ti* __restrict__ cb = a.colval + s4;
        const T* __restrict__ vb = a.nzval + s4;
        const i64 avail = a.nnz_total - s4;  // elements that exist from s4 on
        const int head = (int)(s - s4);      // entries of the previous tile in front of this tile's first one
        for (int base = 0; base < (int)n; base += CHUNK) {
            Ti c[GROUPS][4];
            T v[GROUPS][4];
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const int i = base + (g * THREADS + tid) * 4;
                if (i + 4 <= avail && i < n) {
                    ld4_stream(cb + i, c[g]);
                    ld4_stream(vb + i, v[g]);
                    // the 4-aligned hull [s4, ..) may hold up to 3 entries of the neighbouring tiles on either side: their
                    // columns were not classified with this tile (an interior tile could read a neighbour's ghost column
                    // through the ghost-free view), so they are replaced by a column that is always valid
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (i + k < head || i + k >= (int)n) c[g][k] = (Ti)a.safe_col;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const bool ok = (i + k < avail) && (i + k < n) && (i + k >= head);
                        c[g][k] = ok ? ld_stream(cb + i + k) : (Ti)a.safe_col;
                        v[g][k] = ok ? ld_stream(vb + i + k) : el_zero(T());
                    }
                }
            }
            T xg[GROUPS][4];
#pragma unroll
            for (int g = 0; g < GROUPS; ++g)
#pragma unroll
                for (int k = 0; k < 4; ++k) xg[g][k] = x_at<GHOST, T, Ti>(a.xv, c[g][k]);
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const int i = base + (g * THREADS + tid) * 4;
                if (i < n) {
                    T p[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) p[k] = el_mul(v[g][k], xg[g][k]);
                    st4_shared(prod + i, p);
                }
            }
        }
        __syncthreads();
        // ---- reduce: lanes per row from the tile's mean row length (warp-uniform) ----
        const i64 nrows_t = r1 - r0;
        const i64 avg = (e - s) / nrows_t;
        if (avg <= 12) reduce_rows<T, Ti, THREADS, 1>(prod, a.rowptr, a.y, r0, r1, s4, tid);
        else if (avg <= 24) reduce_rows<T, Ti, THREADS, 2>(prod, a.rowptr, a.y, r0, r1, s4, tid);
        else if (avg <= 48) reduce_rows<T, Ti, THREADS, 4>(prod, a.rowptr, a.y, r0, r1, s4, tid);
        else if (avg <= 96) reduce_rows<T, Ti, THREADS, 8>(prod, a.rowptr, a.y, r0, r1, s4, tid);
        else if (avg <= 192) reduce_rows<T, Ti, THREADS, 16>(prod, a.rowptr, a.y, r0, r1, s4, tid);
        else reduce_rows<T, Ti, THREADS, 32>(prod, a.rowptr, a.y, r0, r1, s4, tid);
    } else {
        // ---- the tile holds a row too long to stage: warp-per-row straight from global memory ----
        const int warp = tid >> 5, lane = tid & 31;
        for (i64 r = r0 + warp; r < r1; r += THREADS / 32) {
            const i64 b = (i64)__ldg(a.rowptr + r) - 1, en = (i64)__ldg(a.rowptr + r + 1) - 1;
            if (en - b > a.long_threshold) continue;  // left to the split kernels
            T acc = el_zero(T());
            for (i64 k = b + lane; k < en; k += 32) {
                const Ti c = ld_stream(a.colval + k);
                acc = el_add(acc, el_mul(ld_stream(a.nzval + k), x_at<GHOST, T, Ti>(a.xv, c)));
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) acc = el_add(acc, shfl_xor(acc, m));
            if (lane == 0) st_y(a.y + r, acc);
        }
    }
}

template <class T, class Ti, int G, bool GHOST, bool DOT = false>
__global__ void __launch_bounds__(ROW_THREADS, RowCfg<T>::CTAS - (GHOST ? 1 : 0))
    spmv_rowwalk_kernel(const TileArgs<T, Ti> a, int cap, int rp_cap, i64 rowptr_len) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Staged<T, Ti> st = stage_tile<T, Ti>(stage_args(a), cap, rp_cap, rowptr_len, smem_raw);
    const double dot = rows_walk<T, Ti, ROW_THREADS, G, GHOST, DOT>(st.scol, st.sval, st.srp, st.rp0, a.xv, a.y, st.r0, st.r1, st.s4, (int)threadIdx.x, a.dot_x);
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

// out2[0] = sum of n per-CTA partials, fixed order (thread t takes t, t + 1024, ...; then a fixed tree); out2[1] = 0
__global__ void __launch_bounds__(1024) dot_partials_sum_kernel(const double* __restrict__ partials, i64 n, double* __restrict__ out2) {
    __shared__ double sh[32];
    double v = 0.0;
    for (i64 i = threadIdx.x; i < n; i += 1024) v += partials[i];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = sh[threadIdx.x];
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (threadIdx.x == 0) out2[0] = v, out2[1] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row walk, direct form: NO dependent load in front of any copy.  A launch is a handful of runs of consecutive tiles
// (TileRuns, by value), so the CTA's tile — hence its window of the nonzero stream — is arithmetic; the per-tile
// HEADER (first row, row count, staged length and the row offsets as 16-bit values relative to the window: a
// library-owned, re-blocked copy of the row-pointer slice at a fixed stride) is arithmetic too.  Thread 0 issues
// everything at once on two mbarriers: A = column indices + header, B = values.  When A lands the lanes issue their x
// gathers; those fly while the values are still landing; when B lands only shared-memory reads and the arithmetic are
// left.  The (rare) rows that overrun the unconditional over-fetch get a second, conditional copy.
// ------------------------------------------------------------------------------------------------------------------
template <class T, class Ti>
struct DirectArgs {
    const Ti* colval;
    const T* nzval;
    const unsigned char* hdrs;  // per tile, hdr_bytes apart: {i64 r0; i32 nrows; i32 n_st; u16 off[nrows+1]}
    XView<T> xv;
    T* y;
    i64 nnz_total;
    TileRuns runs;
    int window, ovf, hdr_bytes;
};

#ifndef HPCLA_DIRECT_CTAS_DROP
#define HPCLA_DIRECT_CTAS_DROP 0  // A/B knob: allocate registers for this many fewer resident CTAs
#endif
#ifndef HPCLA_DIRECT_ONE_BARRIER
#define HPCLA_DIRECT_ONE_BARRIER 0  // A/B knob: values on the same barrier as the column indices (no early gathers)
#endif
template <class T, class Ti, int G>
__global__ void __launch_bounds__(ROW_THREADS, RowCfg<T>::CTAS - HPCLA_DIRECT_CTAS_DROP) spmv_rowwalk_direct_kernel(const DirectArgs<T, Ti> a, int cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* barB = barA + 1;
    const unsigned char* shdr = smem_raw + 16;
    Ti* scol = reinterpret_cast<Ti*>(smem_raw + 16 + a.hdr_bytes);
    T* sval = reinterpret_cast<T*>(scol + cap);
    const int tid = threadIdx.x;
    const i64 tile = tile_of_cta(a.runs);
    const i64 w0 = tile * (i64)a.window;
    const i64 left = (a.nnz_total - w0) & ~(i64)3;
    const int want = a.window + a.ovf;
    const int n_fetch = (int)(left < (i64)want ? (left > 0 ? left : 0) : (i64)want);  // multiple of 4
    uint64_t pol = 0;
    if (tid == 0) {
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        mbar_fence_init();
        pol = l2_evict_first_policy();
        mbar_expect_tx(barA, (uint32_t)n_fetch * (uint32_t)sizeof(Ti) + (uint32_t)a.hdr_bytes);
        bulk_g2s(const_cast<unsigned char*>(shdr), a.hdrs + tile * (i64)a.hdr_bytes, (uint32_t)a.hdr_bytes, barA, pol);
        if (n_fetch > 0) bulk_g2s(scol, a.colval + w0, (uint32_t)n_fetch * (uint32_t)sizeof(Ti), barA, pol);
        mbar_expect_tx(barB, (uint32_t)n_fetch * (uint32_t)sizeof(T));
        if (n_fetch > 0) bulk_g2s(sval, a.nzval + w0, (uint32_t)n_fetch * (uint32_t)sizeof(T), barB, pol);
    }
    __syncthreads();  // barriers initialised before anyone waits
#if HPCLA_DIRECT_ONE_BARRIER
    mbar_wait(barB, 0);
#endif
    mbar_wait(barA, 0);
    const i64 r0 = *reinterpret_cast<const i64*>(shdr);
    const int nrows = *reinterpret_cast<const int*>(shdr + 8);
    const int n_st = *reinterpret_cast<const int*>(shdr + 12);
    const unsigned short* off = reinterpret_cast<const unsigned short*>(shdr + 16);
    if (n_st > n_fetch) {  // uniform, rare: a last row that overruns the over-fetch, or the unaligned tail of the arrays
        const i64 avail = (a.nnz_total - w0 - n_fetch) & ~(i64)3;
        i64 more = ((i64)(n_st - n_fetch) + 3) & ~(i64)3;
        if (more > avail) more = avail;
        __syncthreads();  // everybody has passed the phase-0 wait of barA before it is re-armed
        if (tid == 0) {
            mbar_expect_tx(barA, (uint32_t)more * (uint32_t)(sizeof(Ti) + sizeof(T)));
            if (more > 0) {
                bulk_g2s(scol + n_fetch, a.colval + w0 + n_fetch, (uint32_t)more * (uint32_t)sizeof(Ti), barA, pol);
                bulk_g2s(sval + n_fetch, a.nzval + w0 + n_fetch, (uint32_t)more * (uint32_t)sizeof(T), barA, pol);
            }
        }
        for (int k = n_fetch + (int)more + tid; k < n_st; k += ROW_THREADS) {
            scol[k] = a.colval[w0 + k];
            sval[k] = a.nzval[w0 + k];
        }
        __syncthreads();
        mbar_wait(barA, 1);
    }
    constexpr int RPP = ROW_THREADS / G;
    constexpr int B = HPCLA_WALK_BATCH;
    const int lane = tid % G;
    bool values_in = false;
    for (int base = 0; base < nrows; base += RPP) {
        const int i = base + tid / G;
        const bool valid = i < nrows;
        const int b = valid ? (int)off[i] : 0;
        const int e = valid ? (int)off[i + 1] : 0;
        int k = b + lane;
        T xg[B];  // the gathers go out as soon as the columns are in ...
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int kk = k + u * G;
            xg[u] = el_zero(T());
            if (kk < e) xg[u] = ld_x_pinned(a.xv.own + (i64)scol[kk]);
        }
        if (!values_in) {  // ... and fly while the values are still landing
            mbar_wait(barB, 0);
            values_in = true;
        }
        T acc = el_zero(T());
#pragma unroll
        for (int u = 0; u < B; ++u)
            if (k + u * G < e) acc = el_add(acc, el_mul(sval[k + u * G], xg[u]));
        for (k += B * G; k < e; k += B * G) {  // rows longer than B * G entries
            T p[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int kk = k + u * G;
                p[u] = (kk < e) ? el_mul(sval[kk], x_at<false, T, Ti>(a.xv, scol[kk])) : el_zero(T());
            }
#pragma unroll
            for (int u = 0; u < B; ++u)
                if (k + u * G < e) acc = el_add(acc, p[u]);
        }
        if (G > 1) {
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = el_add(acc, shfl_xor(acc, m));
        }
        if (valid && lane == 0) st_y(a.y + r0 + i, acc);
    }
    if (!values_in) mbar_wait(barB, 0);  // never leave with a copy in flight
}

// Tile headers of the direct row walk (class-1 tiles only): one warp per tile.
template <class Ti>
__global__ void __launch_bounds__(256) build_tile_headers_kernel(const Ti* __restrict__ rowptr, const TileDesc* __restrict__ tiles,
                                                                 const unsigned char* __restrict__ cls, i64 ntiles, int window, int hdr_bytes,
                                                                 unsigned char* __restrict__ hdrs) {
    const i64 t = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles || cls[t] != 1) return;
    unsigned char* h = hdrs + t * (i64)hdr_bytes;
    const i64 r0 = tiles[t].row, r1 = tiles[t + 1].row, e = tiles[t + 1].nnz, w0 = t * (i64)window;
    if (lane == 0) {
        *reinterpret_cast<i64*>(h) = r0;
        *reinterpret_cast<int*>(h + 8) = (int)(r1 - r0);
        *reinterpret_cast<int*>(h + 12) = (int)(e - w0);
    }
    unsigned short* off = reinterpret_cast<unsigned short*>(h + 16);
    for (i64 i = lane; i <= r1 - r0; i += 32) off[i] = (unsigned short)((i64)rowptr[r0 + i] - 1 - w0);
}

// ------------------------------------------------------------------------------------------------------------------
// very long rows ("merge-path split"): the row's nonzero range is cut into equal chunks, one CTA per chunk writes
// one partial sum, a second kernel adds the partials of a row in chunk order.  Deterministic, no atomics.
// ------------------------------------------------------------------------------------------------------------------
template <class T, class Ti, bool GHOST>
__global__ void __launch_bounds__(256) long_rows_partial_kernel(const Ti* __restrict__ rowptr, const Ti* __restrict__ colval,
                                                                const T* __restrict__ nzval, const i64* __restrict__ long_rows,
                                                                const i64* __restrict__ chunk_ptr, i64 nlong, i64 chunk_nnz, XView<T> xv,
                                                                T* __restrict__ partials) {
    __shared__ T sh[32];
    const i64 chunk = blockIdx.x;
    i64 lo = 0, hi = nlong;  // last i with chunk_ptr[i] <= chunk
    while (hi - lo > 1) {
        const i64 mid = (lo + hi) >> 1;
        if (chunk_ptr[mid] <= chunk) lo = mid;
        else hi = mid;
    }
    const i64 r = long_rows[lo];
    const i64 rb = (i64)rowptr[r] - 1, re = (i64)rowptr[r + 1] - 1;
    const i64 b = rb + (chunk - chunk_ptr[lo]) * chunk_nnz;
    const i64 e = (b + chunk_nnz < re) ? b + chunk_nnz : re;
    T acc0 = el_zero(T()), acc1 = el_zero(T());
    i64 k = b + threadIdx.x;
    for (; k + 256 < e; k += 512) {
        const Ti c0 = ld_stream(colval + k), c1 = ld_stream(colval + k + 256);
        const T v0 = ld_stream(nzval + k), v1 = ld_stream(nzval + k + 256);
        acc0 = el_add(acc0, el_mul(v0, x_at<GHOST, T, Ti>(xv, c0)));
        acc1 = el_add(acc1, el_mul(v1, x_at<GHOST, T, Ti>(xv, c1)));
    }
    if (k < e) acc0 = el_add(acc0, el_mul(ld_stream(nzval + k), x_at<GHOST, T, Ti>(xv, ld_stream(colval + k))));
    T tot = block_sum(el_add(acc0, acc1), sh);
    if (threadIdx.x == 0) partials[chunk] = tot;
}

template <class T>
__global__ void long_rows_final_kernel(const i64* __restrict__ long_rows, const i64* __restrict__ chunk_ptr, i64 nlong,
                                       const T* __restrict__ partials, T* __restrict__ y) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlong) return;
    T acc = el_zero(T());
    for (i64 c = chunk_ptr[i]; c < chunk_ptr[i + 1]; ++c) acc = el_add(acc, partials[c]);
    y[long_rows[i]] = acc;
}

// ------------------------------------------------------------------------------------------------------------------
// structure kernels (run once per handle / plan)
// ------------------------------------------------------------------------------------------------------------------
template <class Ti>
__global__ void build_tiles_kernel(const Ti* __restrict__ rowptr, i64 nrows, int window, TileDesc* tiles, i64 ntiles) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > ntiles) return;
    i64 r;
    if (k == ntiles) r = nrows;
    else {
        const i64 target = k * (i64)window;  // first row whose first stored entry is at or after `target`
        i64 lo = 0, hi = nrows + 1;
        while (lo < hi) {
            const i64 mid = (lo + hi) >> 1;
            if ((i64)rowptr[mid] - 1 < target) lo = mid + 1;
            else hi = mid;
        }
        r = lo > nrows ? nrows : lo;
    }
    tiles[k].row = r;
    tiles[k].nnz = (i64)rowptr[r] - 1;
}

// histogram of the row lengths (lengths >= 1023 share the last bin): the most common length sizes the tile window
template <class Ti>
__global__ void __launch_bounds__(256) row_len_hist_kernel(const Ti* __restrict__ rowptr, i64 nrows, unsigned long long* hist /* [1024] */) {
    __shared__ unsigned int sh[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = 0;
    __syncthreads();
    for (i64 r = (i64)blockIdx.x * 256 + threadIdx.x; r < nrows; r += (i64)gridDim.x * 256) {
        const i64 len = (i64)rowptr[r + 1] - (i64)rowptr[r];
        atomicAdd(&sh[len < 1023 ? (int)len : 1023], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += 256)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// one-time check of a borrowed CSR view (hpcla_csr_create): a bad row pointer or column would otherwise index shared
// memory with a negative length or read x out of bounds on the device
template <class Ti>
__global__ void validate_csr_kernel(const Ti* __restrict__ rowptr, const Ti* __restrict__ colval, i64 nrows, i64 nnz, i64 ncc, unsigned* bad) {
    const i64 stride = (i64)gridDim.x * blockDim.x, t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned mine = 0;
    if (t == 0) mine += ((i64)rowptr[0] != 1) + ((i64)rowptr[nrows] != nnz + 1);
    for (i64 r = t; r < nrows; r += stride) mine += (i64)rowptr[r] > (i64)rowptr[r + 1];
    for (i64 k = t; k < nnz; k += stride) mine += (unsigned long long)((i64)colval[k] - 1) >= (unsigned long long)ncc;
    if (mine) atomicAdd(bad, mine);
}

template <class Ti>
__global__ void find_long_rows_kernel(const Ti* __restrict__ rowptr, i64 nrows, i64 threshold, i64* rows_out, i64 cap,
                                      unsigned long long* count) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    if ((i64)rowptr[r + 1] - (i64)rowptr[r] > threshold) {
        const unsigned long long slot = atomicAdd(count, 1ull);
        if ((i64)slot < cap) rows_out[slot] = r;
    }
}

template <class Ti>
__global__ void __launch_bounds__(256) classify_tiles_kernel(const Ti* __restrict__ colval, const TileDesc* __restrict__ tiles, i64 own_lo,
                                                             unsigned long long own_n, unsigned char* flags) {
    const i64 t = blockIdx.x;
    const i64 b = tiles[t].nnz, e = tiles[t + 1].nnz;
    int ghost = 0;
    for (i64 k = b + threadIdx.x; k < e; k += 256) ghost |= ((unsigned long long)((i64)colval[k] - own_lo) >= own_n);
    ghost = __syncthreads_or(ghost);
    if (threadIdx.x == 0) flags[t] = ghost ? 1 : 0;
}

// Kernel class of every tile, one warp per tile: 0 = no row starts in the tile's window, 1 = row walk (all rows
// completely staged within `cap` nonzeros and `rp_cap` row pointers, and mean row length >= half the longest, i.e. the
// lanes of the walk are well used), 2 = general kernel.
template <class Ti>
__global__ void __launch_bounds__(256) tile_class_kernel(const Ti* __restrict__ rowptr, const TileDesc* __restrict__ tiles, i64 ntiles, int window, int cap,
                                                         int rp_cap, int balance_pct, unsigned char* cls) {
    const i64 t = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const i64 r0 = tiles[t].row, r1 = tiles[t + 1].row, s = tiles[t].nnz, e = tiles[t + 1].nnz;
    if (r1 <= r0) {
        if (lane == 0) cls[t] = 0;
        return;
    }
    i64 maxlen = 0;
    for (i64 r = r0 + lane; r < r1; r += 32) {
        const i64 len = (i64)rowptr[r + 1] - (i64)rowptr[r];
        maxlen = len > maxlen ? len : maxlen;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const i64 o = __shfl_xor_sync(0xffffffffu, maxlen, m);
        maxlen = o > maxlen ? o : maxlen;
    }
    // staged from the window start; rp_cap - 8 rows is also what a tile header of the direct walk holds
    const bool fits = (e - t * (i64)window) <= (i64)cap && (r1 + 1 - (r0 & ~(i64)3)) <= (i64)rp_cap && (r1 - r0) <= (i64)(rp_cap - 8);
    const bool balanced = 100 * (e - s) >= (i64)balance_pct * (r1 - r0) * maxlen;  // mean row length >= balance_pct % of the longest
    if (lane == 0) cls[t] = (fits && balanced && e > s) ? 1 : 2;
}

// Largest own column (0-based, relative to the own segment) referenced by each tile, -1 if none: the prefix of x.v a
// tile needs before it can run (staged host -> device pipeline, hpcla_spmv_run_staged).
template <class Ti>
__global__ void __launch_bounds__(256) tile_maxcol_kernel(const Ti* __restrict__ colval, const TileDesc* __restrict__ tiles, i64 own_lo,
                                                          unsigned long long own_n, i64* __restrict__ out) {
    __shared__ i64 sh[8];
    const i64 t = blockIdx.x;
    const i64 b = tiles[t].nnz, e = tiles[t + 1].nnz;
    i64 m = -1;
    for (i64 k = b + threadIdx.x; k < e; k += 256) {
        const i64 c = (i64)colval[k] - own_lo;
        if ((unsigned long long)c < own_n && c > m) m = c;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const i64 o = __shfl_xor_sync(0xffffffffu, m, d);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = sh[w] > m ? sh[w] : m;
        out[t] = m;
    }
}

template <class T> __global__ void pack_kernel(const T* __restrict__ x, const i64* __restrict__ idx, i64 n, T* __restrict__ out);
template <class T> __global__ void local_copy_kernel(const T* __restrict__ x, const i64* __restrict__ src, const i64* __restrict__ dst, i64 n, T* __restrict__ g);

// Direct halo: raise a flag in (possibly peer) memory once everything before it in the stream — the copy-engine push of
// the ghost values — is complete.  System-scope release: the flag may be polled by another GPU's front end.
__global__ void write_flag_kernel(unsigned* flag, unsigned value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
// Loads (CUDA loads kernels lazily, on first launch, and a load synchronises the context) every kernel an exchange may
// launch for the first time while some stream already sits in a flag wait: between the rank-threads of one process such a
// load would wait for a halo stream that waits for a push this very thread has not enqueued yet.
cudaError_t preload_halo_kernels() {
    cudaFuncAttributes a;
    cudaError_t e;
    if ((e = cudaFuncGetAttributes(&a, write_flag_kernel)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, pack_kernel<float>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, pack_kernel<double>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, pack_kernel<double2>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, local_copy_kernel<float>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, local_copy_kernel<double>)) != cudaSuccess) return e;
    if ((e = cudaFuncGetAttributes(&a, local_copy_kernel<double2>)) != cudaSuccess) return e;
    return cudaSuccess;
}
cudaError_t launch_write_flag(unsigned* flag, unsigned value, cudaStream_t st) {
    write_flag_kernel<<<1, 1, 0, st>>>(flag, value);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// halo pack / local copy
// ------------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void pack_kernel(const T* __restrict__ x, const i64* __restrict__ idx, i64 n, T* __restrict__ out) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = x[idx[k] - 1];
}
template <class T>
__global__ void local_copy_kernel(const T* __restrict__ x, const i64* __restrict__ src, const i64* __restrict__ dst, i64 n, T* __restrict__ g) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) g[dst[k] - 1] = x[src[k] - 1];
}

// ------------------------------------------------------------------------------------------------------------------
// reductions and vector updates (HPCVector dot / norm / broadcast, src/vectors.jl:758-812, 1203-1221)
// Two-level deterministic sum: per-CTA partials, the last CTA to finish adds them in index order.
// ------------------------------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 148 * 8;

struct d2 {
    double re, im;
};
__device__ __forceinline__ void acc_dot(d2& a, float x, float y) { a.re += (double)x * (double)y; }
__device__ __forceinline__ void acc_dot(d2& a, double x, double y) { a.re += x * y; }
__device__ __forceinline__ void acc_dot(d2& a, cplx x, cplx y) {  // conj(x) * y
    a.re += x.re * y.re + x.im * y.im;
    a.im += x.re * y.im - x.im * y.re;
}

// scratch layout: [0 .. 2*RED_MAX_BLOCKS) partial (re, im) pairs, then one unsigned counter (as a double slot)
__device__ __forceinline__ void finish_reduction(d2 mine, double* scratch, double* out2) {
    __shared__ double shre[32], shim[32];
    __shared__ bool is_last;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        mine.re += __shfl_xor_sync(0xffffffffu, mine.re, m);
        mine.im += __shfl_xor_sync(0xffffffffu, mine.im, m);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) shre[warp] = mine.re, shim[warp] = mine.im;
    __syncthreads();
    if (threadIdx.x == 0) {
        double re = 0, im = 0;
        for (int w = 0; w < RED_THREADS / 32; ++w) re += shre[w], im += shim[w];
        scratch[2 * blockIdx.x] = re;
        scratch[2 * blockIdx.x + 1] = im;
        __threadfence();
        unsigned* counter = reinterpret_cast<unsigned*>(scratch + 2 * RED_MAX_BLOCKS);
        const unsigned done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double re = 0, im = 0;  // fixed order: thread t takes partials t, t+256, ...; then the same tree as above
        for (int b = threadIdx.x; b < (int)gridDim.x; b += RED_THREADS) {
            re += __ldcg(scratch + 2 * b);
            im += __ldcg(scratch + 2 * b + 1);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            re += __shfl_xor_sync(0xffffffffu, re, m);
            im += __shfl_xor_sync(0xffffffffu, im, m);
        }
        __syncthreads();
        if (lane == 0) shre[warp] = re, shim[warp] = im;
        __syncthreads();
        if (threadIdx.x == 0) {
            double r2 = 0, i2 = 0;
            for (int w = 0; w < RED_THREADS / 32; ++w) r2 += shre[w], i2 += shim[w];
            out2[0] = r2;
            out2[1] = i2;
            *reinterpret_cast<unsigned*>(scratch + 2 * RED_MAX_BLOCKS) = 0u;  // re-arm
        }
    }
}

template <class T>
__global__ void __launch_bounds__(RED_THREADS) dot_kernel(i64 n, const T* __restrict__ x, const T* __restrict__ y, double* scratch, double* out2) {
    d2 a{0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * RED_THREADS) acc_dot(a, ld_x(x + i), ld_x(y + i));
    finish_reduction(a, scratch, out2);
}

template <class T>
struct Scal {
    T v;
};
__device__ __forceinline__ float axpby1(float a, float x, float b, float y) { return a * x + b * y; }
__device__ __forceinline__ double axpby1(double a, double x, double b, double y) { return a * x + b * y; }
__device__ __forceinline__ cplx axpby1(cplx a, cplx x, cplx b, cplx y) { return el_add(el_mul(a, x), el_mul(b, y)); }
template <class T>
__global__ void axpby_kernel(i64 n, Scal<T> alpha, const T* __restrict__ x, Scal<T> beta, T* __restrict__ y) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) y[i] = axpby1(alpha.v, x[i], beta.v, y[i]);
}

// CG: r = b, p = b, x = 0, rr = dot(b, b)
template <class T>
__global__ void __launch_bounds__(RED_THREADS) cg_init_kernel(i64 n, const T* __restrict__ b, T* __restrict__ x, T* __restrict__ r, T* __restrict__ p,
                                                              double* scratch, double* rr_out2) {
    d2 a{0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * RED_THREADS) {
        const T bi = b[i];
        x[i] = (T)0;
        r[i] = bi;
        p[i] = bi;
        acc_dot(a, bi, bi);
    }
    finish_reduction(a, scratch, rr_out2);
}
// alpha = rr/pq ; x += alpha p ; r -= alpha q ; rr_new = dot(r, r)
template <class T>
__global__ void __launch_bounds__(RED_THREADS) cg_update_xr_kernel(i64 n, const T* __restrict__ p, const T* __restrict__ q, T* __restrict__ x,
                                                                   T* __restrict__ r, const double* __restrict__ rr, const double* __restrict__ pq,
                                                                   double* scratch, double* rr_new_out2) {
    const T alpha = (T)(rr[0] / pq[0]);
    d2 a{0.0, 0.0};
    for (i64 i = (i64)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (i64)gridDim.x * RED_THREADS) {
        x[i] = x[i] + alpha * p[i];
        const T ri = r[i] - alpha * q[i];
        r[i] = ri;
        acc_dot(a, ri, ri);
    }
    finish_reduction(a, scratch, rr_new_out2);
}
// beta = rr_new/rr ; p = r + beta p
template <class T>
__global__ void cg_update_p_kernel(i64 n, const T* __restrict__ r, T* __restrict__ p, const double* __restrict__ rr_new, const double* __restrict__ rr) {
    const T beta = (T)(rr_new[0] / rr[0]);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) p[i] = r[i] + beta * p[i];
}

// ==================================================================================================================
// host-side launchers
// ==================================================================================================================

// Tile shape of one matrix.  Row walk: G lanes per row and a window of about (256 / G) mean-length rows, G the
// smallest power of two for which a tile's operands stay within the shared-memory budget; the general kernel stages
// products for the same tiles.  `irregular` (most nonzeros in tiles the row walk cannot take): the general kernel's
// own shape.
template <class T>
static TileShape shape_of(int itype, double avg_row, bool irregular, int lanes_override, int window_override) {
    TileShape s;
    s.threads = ROW_THREADS;
    const int chunk = TileCfg<T>::THREADS * TileCfg<T>::GROUPS * 4;
    const size_t per_nz = sizeof(T) + (itype == HPCLA_I32 ? 4 : 8);
    if (irregular) {
        s.lanes = 0;
        s.window = chunk - 64;
        s.cap = 0;
        s.rp_cap = 0;
        s.ovf = s.hdr_rows = s.hdr_bytes = 0;
        s.general_elems = chunk + 512;
    } else {
        // measured on B200 (profiles/r1d_tune_*.txt): 27-point rows want ~70 KB tiles (G = 2: 128 rows, 3 CTAs per SM),
        // 7-point rows fit one row per thread in 21 KB (7 CTAs per SM)
        const double budget = 72.0 * 1024;
        int G = 1;
        while (G < 32 && (ROW_THREADS / G) * avg_row * (double)per_nz > budget) G *= 2;
        if (lanes_override == 1 || lanes_override == 2 || lanes_override == 4 || lanes_override == 8 || lanes_override == 16 || lanes_override == 32)
            G = lanes_override;
        s.lanes = G;
        double w = (ROW_THREADS / G) * (avg_row > 1.0 ? avg_row : 1.0);
        const double wmax = budget / (double)per_nz;
        if (w > wmax) w = wmax;
        s.window = ((int)w + 3) & ~3;
        if (s.window < 256) s.window = 256;
        if (window_override >= 64) s.window = window_override & ~3;
        int slack = ((int)(4 * avg_row) + 31) & ~31;
        slack = slack < 64 ? 64 : slack > 512 ? 512 : slack;
        s.cap = s.window + slack;
        // the direct walk over-fetches `ovf` entries past the window unconditionally (a typical last row never needs a
        // second copy) and holds up to hdr_rows 16-bit row offsets per tile header
        s.ovf = (((int)avg_row + 4) + 3) & ~3;
        if (s.ovf > slack) s.ovf = slack;
        s.hdr_rows = 2 * (ROW_THREADS / G);  // boundary rows of a stencil are shorter: up to twice the typical row count
        const int by_window = (int)(2.0 * s.window / (avg_row > 1.0 ? avg_row : 1.0));  // (a window override may hold more rows)
        if (by_window > s.hdr_rows) s.hdr_rows = (by_window + 7) & ~7;
        if (s.cap > 65535) s.hdr_rows = 0;  // offsets would not fit 16 bits: no direct walk
        s.hdr_bytes = (16 + 2 * (s.hdr_rows + 1) + 15) & ~15;
        s.rp_cap = (s.hdr_rows + 8 + 3) & ~3;
        s.general_elems = s.cap + 256;
    }
    if (window_override >= 64 && irregular && window_override <= s.general_elems - 8) s.window = window_override & ~3;
    return s;
}
TileShape tile_shape(int dtype, int itype, double avg_row, bool irregular, int lanes_override, int window_override) {
    if (dtype == HPCLA_F32) return shape_of<float>(itype, avg_row, irregular, lanes_override, window_override);
    if (dtype == HPCLA_F64) return shape_of<double>(itype, avg_row, irregular, lanes_override, window_override);
    return shape_of<cplx>(itype, avg_row, irregular, lanes_override, window_override);
}

static inline int blocks_for(i64 n, int threads) { return (int)((n + threads - 1) / threads); }

cudaError_t launch_build_tiles(int itype, const void* rowptr, i64 nrows, i64 nnz, int window, TileDesc* tiles, i64 ntiles, cudaStream_t st) {
    (void)nnz;
    const int blocks = blocks_for(ntiles + 1, 256);
    if (itype == HPCLA_I32) build_tiles_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, nrows, window, tiles, ntiles);
    else build_tiles_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, nrows, window, tiles, ntiles);
    return cudaGetLastError();
}

cudaError_t launch_tile_class(int itype, const void* rowptr, const TileDesc* tiles, i64 ntiles, int window, int cap, int rp_cap, int balance_pct,
                              unsigned char* cls, cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    const int blocks = blocks_for(ntiles, 8);
    if (itype == HPCLA_I32) tile_class_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, tiles, ntiles, window, cap, rp_cap, balance_pct, cls);
    else tile_class_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, tiles, ntiles, window, cap, rp_cap, balance_pct, cls);
    return cudaGetLastError();
}

cudaError_t launch_row_len_hist(int itype, const void* rowptr, i64 nrows, unsigned long long* hist, cudaStream_t st) {
    if (nrows == 0) return cudaSuccess;
    const int blocks = (int)std::min<i64>(148 * 8, (nrows + 255) / 256);
    if (itype == HPCLA_I32) row_len_hist_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, nrows, hist);
    else row_len_hist_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, nrows, hist);
    return cudaGetLastError();
}

cudaError_t launch_validate_csr(int itype, const void* rowptr, const void* colval, i64 nrows, i64 nnz, i64 ncc, unsigned* d_bad, cudaStream_t st) {
    const int blocks = (int)std::min<i64>(148 * 8, (std::max(nrows, nnz) + 255) / 256 + 1);
    if (itype == HPCLA_I32) validate_csr_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, (const int*)colval, nrows, nnz, ncc, d_bad);
    else validate_csr_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, (const long long*)colval, nrows, nnz, ncc, d_bad);
    return cudaGetLastError();
}

cudaError_t launch_find_long_rows(int itype, const void* rowptr, i64 nrows, i64 threshold, i64* rows_out, i64 cap, unsigned long long* count_out,
                                  cudaStream_t st) {
    if (nrows == 0) return cudaSuccess;
    const int blocks = blocks_for(nrows, 256);
    if (itype == HPCLA_I32) find_long_rows_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, nrows, threshold, rows_out, cap, count_out);
    else find_long_rows_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, nrows, threshold, rows_out, cap, count_out);
    return cudaGetLastError();
}

cudaError_t launch_classify_tiles(int itype, const void* colval, const TileDesc* tiles, i64 ntiles, i64 own_lo, i64 own_n, unsigned char* flags,
                                  cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    if (itype == HPCLA_I32) classify_tiles_kernel<int><<<(unsigned)ntiles, 256, 0, st>>>((const int*)colval, tiles, own_lo, (unsigned long long)own_n, flags);
    else classify_tiles_kernel<long long><<<(unsigned)ntiles, 256, 0, st>>>((const long long*)colval, tiles, own_lo, (unsigned long long)own_n, flags);
    return cudaGetLastError();
}

cudaError_t launch_tile_maxcol(int itype, const void* colval, const TileDesc* tiles, i64 ntiles, i64 own_lo, i64 own_n, i64* out, cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    if (itype == HPCLA_I32) tile_maxcol_kernel<int><<<(unsigned)ntiles, 256, 0, st>>>((const int*)colval, tiles, own_lo, (unsigned long long)own_n, out);
    else tile_maxcol_kernel<long long><<<(unsigned)ntiles, 256, 0, st>>>((const long long*)colval, tiles, own_lo, (unsigned long long)own_n, out);
    return cudaGetLastError();
}

template <class T>
static XView<T> make_xview(const void* x_own, const void* gathered, i64 own_lo, i64 own_n) {
    XView<T> v;
    v.own = x_own ? (const T*)x_own - own_lo : nullptr;
    v.gat = gathered ? (const T*)gathered - 1 : nullptr;
    v.own_lo = own_lo;
    v.own_n = (unsigned long long)own_n;
    return v;
}


template <class T, class Ti>
static TileArgs<T, Ti> tile_args(const SpmvLaunch& L) {
    TileArgs<T, Ti> a;
    a.rowptr = (const Ti*)L.rowptr;
    a.colval = (const Ti*)L.colval;
    a.nzval = (const T*)L.nzval;
    a.xv = make_xview<T>(L.x_own, L.gathered, L.own_lo, L.own_n);
    a.y = (T*)L.y;
    a.recs = L.recs;
    a.runs = launch_runs(L.n_runs, L.run_cta0, L.run_tile0);
    a.window = L.shape.window;
    a.nnz_total = L.nnz;
    a.long_threshold = L.long_threshold;
    a.safe_col = L.own_n > 0 ? L.own_lo : 1;
    a.dot_x = (const T*)L.dot_x;
    a.dot_out = L.dot_out;
    return a;
}

size_t direct_smem_bytes(int dtype, int itype, const TileShape& sh) {
    const size_t ts = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16, is = itype == HPCLA_I32 ? 4 : 8;
    return 16 + (size_t)sh.hdr_bytes + (size_t)sh.cap * (is + ts);
}

size_t rowwalk_smem_bytes(int dtype, int itype, const TileShape& sh) {
    const size_t ts = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16, is = itype == HPCLA_I32 ? 4 : 8;
    return 16 + is * (size_t)sh.rp_cap + (size_t)sh.cap * (is + ts);
}

template <class T, class Ti, int G>
static cudaError_t rowwalk_launch(const SpmvLaunch& L, const TileArgs<T, Ti>& a, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if constexpr (!std::is_same<T, cplx>::value) {
        if (L.dot_x) {  // the instantiations that also leave dot(dot_x, y) partials (real types: CG)
            if (L.has_ghost) {
                if ((e = ensure_smem<spmv_rowwalk_kernel<T, Ti, G, true, true>>(smem, true)) != cudaSuccess) return e;
                spmv_rowwalk_kernel<T, Ti, G, true, true><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
            } else {
                if ((e = ensure_smem<spmv_rowwalk_kernel<T, Ti, G, false, true>>(smem, true)) != cudaSuccess) return e;
                spmv_rowwalk_kernel<T, Ti, G, false, true><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
            }
            return cudaGetLastError();
        }
    }
    if (L.has_ghost) {
        if ((e = ensure_smem<spmv_rowwalk_kernel<T, Ti, G, true>>(smem, true)) != cudaSuccess) return e;
        spmv_rowwalk_kernel<T, Ti, G, true><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
    } else {
        if ((e = ensure_smem<spmv_rowwalk_kernel<T, Ti, G, false>>(smem, true)) != cudaSuccess) return e;
        spmv_rowwalk_kernel<T, Ti, G, false><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
    }
    return cudaGetLastError();
}

template <class T, class Ti>
static cudaError_t spmv_rowwalk_typed(const SpmvLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    const TileArgs<T, Ti> a = tile_args<T, Ti>(L);
    const size_t smem = rowwalk_smem_bytes(L.dtype, L.itype, L.shape);
    switch (L.shape.lanes) {
        case 1: return rowwalk_launch<T, Ti, 1>(L, a, smem, st);
        case 2: return rowwalk_launch<T, Ti, 2>(L, a, smem, st);
        case 4: return rowwalk_launch<T, Ti, 4>(L, a, smem, st);
        case 8: return rowwalk_launch<T, Ti, 8>(L, a, smem, st);
        case 16: return rowwalk_launch<T, Ti, 16>(L, a, smem, st);
        case 32: return rowwalk_launch<T, Ti, 32>(L, a, smem, st);
    }
    return cudaErrorInvalidValue;
}

template <class T, class Ti, int G>
static cudaError_t direct_launch(const SpmvLaunch& L, const DirectArgs<T, Ti>& a, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if ((e = ensure_smem<spmv_rowwalk_direct_kernel<T, Ti, G>>(smem, true)) != cudaSuccess) return e;
    spmv_rowwalk_direct_kernel<T, Ti, G><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap);
    return cudaGetLastError();
}

template <class T, class Ti>
static cudaError_t spmv_direct_typed(const SpmvLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    DirectArgs<T, Ti> a;
    a.colval = (const Ti*)L.colval;
    a.nzval = (const T*)L.nzval;
    a.hdrs = L.hdrs;
    a.xv = make_xview<T>(L.x_own, L.gathered, L.own_lo, L.own_n);
    a.y = (T*)L.y;
    a.nnz_total = L.nnz;
    a.runs = launch_runs(L.n_runs, L.run_cta0, L.run_tile0);
    a.window = L.shape.window;
    a.ovf = L.shape.ovf;
    a.hdr_bytes = L.shape.hdr_bytes;
    const size_t smem = direct_smem_bytes(L.dtype, L.itype, L.shape);
    switch (L.shape.lanes) {
        case 1: return direct_launch<T, Ti, 1>(L, a, smem, st);
        case 2: return direct_launch<T, Ti, 2>(L, a, smem, st);
        case 4: return direct_launch<T, Ti, 4>(L, a, smem, st);
        case 8: return direct_launch<T, Ti, 8>(L, a, smem, st);
        case 16: return direct_launch<T, Ti, 16>(L, a, smem, st);
        case 32: return direct_launch<T, Ti, 32>(L, a, smem, st);
    }
    return cudaErrorInvalidValue;
}

template <class T, class Ti>
static cudaError_t spmv_general_typed(const SpmvLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    constexpr int THREADS = TileCfg<T>::THREADS, GROUPS = TileCfg<T>::GROUPS;
    const TileArgs<T, Ti> a = tile_args<T, Ti>(L);
    const size_t smem = (size_t)L.shape.general_elems * sizeof(T);
    cudaError_t e;
    if (L.has_ghost) {
        if ((e = ensure_smem<spmv_tile_kernel<T, Ti, THREADS, GROUPS, true>>(smem, false)) != cudaSuccess) return e;
        spmv_tile_kernel<T, Ti, THREADS, GROUPS, true><<<L.n_launch, THREADS, smem, st>>>(a, L.shape.general_elems);
    } else {
        if ((e = ensure_smem<spmv_tile_kernel<T, Ti, THREADS, GROUPS, false>>(smem, false)) != cudaSuccess) return e;
        spmv_tile_kernel<T, Ti, THREADS, GROUPS, false><<<L.n_launch, THREADS, smem, st>>>(a, L.shape.general_elems);
    }
    return cudaGetLastError();
}

cudaError_t launch_spmv_rowwalk(const SpmvLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(spmv_rowwalk_typed, L, L, st); }
cudaError_t launch_dot_partials_sum(const double* partials, i64 n, double* out2, cudaStream_t st) {
    dot_partials_sum_kernel<<<1, 1024, 0, st>>>(partials, n, out2);
    return cudaGetLastError();
}
cudaError_t launch_spmv_general(const SpmvLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(spmv_general_typed, L, L, st); }
cudaError_t launch_spmv_direct(const SpmvLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(spmv_direct_typed, L, L, st); }

cudaError_t launch_build_tile_headers(int itype, const void* rowptr, const TileDesc* tiles, const unsigned char* cls, i64 ntiles, int window, int hdr_bytes,
                                      unsigned char* hdrs, cudaStream_t st) {
    if (ntiles == 0) return cudaSuccess;
    const int blocks = blocks_for(ntiles, 8);
    if (itype == HPCLA_I32) build_tile_headers_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, tiles, cls, ntiles, window, hdr_bytes, hdrs);
    else build_tile_headers_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, tiles, cls, ntiles, window, hdr_bytes, hdrs);
    return cudaGetLastError();
}

template <class T, class Ti>
static cudaError_t long_rows_typed(const LongRowsLaunch& L, cudaStream_t st) {
    if (L.nlong == 0) return cudaSuccess;
    XView<T> xv = make_xview<T>(L.x_own, L.gathered, L.own_lo, L.own_n);
    if (L.has_ghost)
        long_rows_partial_kernel<T, Ti, true><<<(unsigned)L.nchunks, 256, 0, st>>>((const Ti*)L.rowptr, (const Ti*)L.colval, (const T*)L.nzval, L.long_rows,
                                                                                  L.chunk_ptr, L.nlong, L.chunk_nnz, xv, (T*)L.partials);
    else
        long_rows_partial_kernel<T, Ti, false><<<(unsigned)L.nchunks, 256, 0, st>>>((const Ti*)L.rowptr, (const Ti*)L.colval, (const T*)L.nzval, L.long_rows,
                                                                                   L.chunk_ptr, L.nlong, L.chunk_nnz, xv, (T*)L.partials);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    long_rows_final_kernel<T><<<blocks_for(L.nlong, 128), 128, 0, st>>>(L.long_rows, L.chunk_ptr, L.nlong, (const T*)L.partials, (T*)L.y);
    return cudaGetLastError();
}
cudaError_t launch_long_rows(const LongRowsLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(long_rows_typed, L, L, st); }

#define HPCLA_DISPATCH_T(dtype, CALL_F32, CALL_F64, CALL_C128) \
    do {                                                       \
        if ((dtype) == HPCLA_F32) { CALL_F32; }                \
        else if ((dtype) == HPCLA_F64) { CALL_F64; }           \
        else if ((dtype) == HPCLA_C128) { CALL_C128; }         \
        else return cudaErrorInvalidValue;                     \
    } while (0)

cudaError_t launch_pack(int dtype, const void* x, const i64* idx, i64 n, void* out, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int b = blocks_for(n, 256);
    HPCLA_DISPATCH_T(dtype, (pack_kernel<float><<<b, 256, 0, st>>>((const float*)x, idx, n, (float*)out)),
                     (pack_kernel<double><<<b, 256, 0, st>>>((const double*)x, idx, n, (double*)out)),
                     (pack_kernel<double2><<<b, 256, 0, st>>>((const double2*)x, idx, n, (double2*)out)));
    return cudaGetLastError();
}
cudaError_t launch_local_copy(int dtype, const void* x, const i64* src, const i64* dst, i64 n, void* g, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const int b = blocks_for(n, 256);
    HPCLA_DISPATCH_T(dtype, (local_copy_kernel<float><<<b, 256, 0, st>>>((const float*)x, src, dst, n, (float*)g)),
                     (local_copy_kernel<double><<<b, 256, 0, st>>>((const double*)x, src, dst, n, (double*)g)),
                     (local_copy_kernel<double2><<<b, 256, 0, st>>>((const double2*)x, src, dst, n, (double2*)g)));
    return cudaGetLastError();
}

int reduce_scratch_doubles() { return 2 * RED_MAX_BLOCKS + 2; }
static inline int red_blocks(i64 n) {
    i64 b = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
    if (b < 1) b = 1;
    if (b > RED_MAX_BLOCKS) b = RED_MAX_BLOCKS;
    return (int)b;
}

cudaError_t launch_dot(int dtype, i64 n, const void* x, const void* y, double* scratch, double* out2, cudaStream_t st) {
    const int b = red_blocks(n);
    HPCLA_DISPATCH_T(dtype, (dot_kernel<float><<<b, RED_THREADS, 0, st>>>(n, (const float*)x, (const float*)y, scratch, out2)),
                     (dot_kernel<double><<<b, RED_THREADS, 0, st>>>(n, (const double*)x, (const double*)y, scratch, out2)),
                     (dot_kernel<cplx><<<b, RED_THREADS, 0, st>>>(n, (const cplx*)x, (const cplx*)y, scratch, out2)));
    return cudaGetLastError();
}
cudaError_t launch_axpby(int dtype, i64 n, const void* alpha, const void* x, const void* beta, void* y, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    i64 bl = (n + 1023) / 1024;
    const int b = (int)(bl > 148 * 16 ? 148 * 16 : bl);
    HPCLA_DISPATCH_T(dtype, (axpby_kernel<float><<<b, 256, 0, st>>>(n, Scal<float>{*(const float*)alpha}, (const float*)x, Scal<float>{*(const float*)beta}, (float*)y)),
                     (axpby_kernel<double><<<b, 256, 0, st>>>(n, Scal<double>{*(const double*)alpha}, (const double*)x, Scal<double>{*(const double*)beta}, (double*)y)),
                     (axpby_kernel<cplx><<<b, 256, 0, st>>>(n, Scal<cplx>{*(const cplx*)alpha}, (const cplx*)x, Scal<cplx>{*(const cplx*)beta}, (cplx*)y)));
    return cudaGetLastError();
}

cudaError_t launch_cg_init(int dtype, i64 n, const void* b, void* x, void* r, void* p, double* scratch, double* rr_out2, cudaStream_t st) {
    const int g = red_blocks(n);
    if (dtype == HPCLA_F32) cg_init_kernel<float><<<g, RED_THREADS, 0, st>>>(n, (const float*)b, (float*)x, (float*)r, (float*)p, scratch, rr_out2);
    else if (dtype == HPCLA_F64) cg_init_kernel<double><<<g, RED_THREADS, 0, st>>>(n, (const double*)b, (double*)x, (double*)r, (double*)p, scratch, rr_out2);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_cg_update_xr(int dtype, i64 n, const void* p, const void* q, void* x, void* r, const double* rr, const double* pq, double* scratch,
                                double* rr_new_out2, cudaStream_t st) {
    const int g = red_blocks(n);
    if (dtype == HPCLA_F32)
        cg_update_xr_kernel<float><<<g, RED_THREADS, 0, st>>>(n, (const float*)p, (const float*)q, (float*)x, (float*)r, rr, pq, scratch, rr_new_out2);
    else if (dtype == HPCLA_F64)
        cg_update_xr_kernel<double><<<g, RED_THREADS, 0, st>>>(n, (const double*)p, (const double*)q, (double*)x, (double*)r, rr, pq, scratch, rr_new_out2);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
cudaError_t launch_cg_update_p(int dtype, i64 n, const void* r, void* p, const double* rr_new, const double* rr, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    i64 bl = (n + 1023) / 1024;
    const int g = (int)(bl > 148 * 16 ? 148 * 16 : bl);
    if (dtype == HPCLA_F32) cg_update_p_kernel<float><<<g, 256, 0, st>>>(n, (const float*)r, (float*)p, rr_new, rr);
    else if (dtype == HPCLA_F64) cg_update_p_kernel<double><<<g, 256, 0, st>>>(n, (const double*)r, (double*)p, rr_new, rr);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace hpcla
