// flat.cu — the multiply for IRREGULAR matrices (power-law row lengths, BASELINE config 4): an nnz-split
// ("merge-path") kernel.  Replaces the reference's one-work-item-per-row _spmv_kernel! (src/sparse.jl:2055-2066) where
// row lengths spread over five orders of magnitude and a row-per-thread (or row-per-tile) assignment cannot balance.
//
// The stored entries are cut into CHUNKS of 4096 consecutive nonzeros, one CTA each, 512 per warp, regardless of where
// rows begin or end:
//   * the chunk's slices of colval / nzval and its 512 bytes of ROW-START FLAGS (one bit per stored entry, built once
//     per matrix) are fetched with bulk asynchronous copies (cp.async.bulk, the 1-D TMA path: the matrix stream never
//     touches the LSU / L1TEX pipe, which is what the scattered x gathers are short of);
//   * every lane owns 4 consecutive entries per round (conflict-free 128-bit shared-memory reads), gathers x for all its
//     16 entries up front, and the warp reduces by row with a register-level segmented scan (shuffles + ballots): no
//     CTA-wide barrier after the copy has landed, no products staged in shared memory;
//   * rows that end inside the warp's 512 entries are written to y directly; the partial sum in front of the first row
//     start of a warp chunk goes to heads[chunk], and a small second kernel adds the heads of the following chunks to
//     the row they continue, in chunk order (deterministic, no atomics).
// Rows above the split threshold are recomputed by the long-row kernels afterwards, as before.
// Row pointers are not read at all per multiply: the flags carry the row structure (1 bit per entry instead of
// sizeof(Ti) per row).
#include <cub/device/device_scan.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "device_common.cuh"

namespace hpcla {

constexpr int FLAT_THREADS = 256;
constexpr int FLAT_ROUNDS = 4;                               // rounds of 4 entries per lane
constexpr int FLAT_WCHUNK = 32 * 4 * FLAT_ROUNDS;            // 512 entries per warp
constexpr int FLAT_CHUNK = (FLAT_THREADS / 32) * FLAT_WCHUNK;  // 4096 entries per CTA
constexpr int FLAT_WORDS = FLAT_CHUNK / 32;                  // flag words per CTA chunk

template <class T, class Ti>
struct FlatArgs {
    const Ti* colval;
    const T* nzval;
    const unsigned* bits;  // row-start flags, FLAT_WORDS words per chunk
    const i64* wrow;       // per warp chunk: ordinal (among the non-empty rows) of the row holding its first entry
    const i64* row_map;    // ordinal -> local row, or null when no row is empty
    XView<T> xv;
    T* y;
    T* heads;  // per warp chunk: the partial sum in front of its first row start
    i64 nnz;
    i64 safe_col;
};

__device__ __forceinline__ float shfl_up(float v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_up(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ cplx shfl_up(cplx v, int d) { return cplx{__shfl_up_sync(0xffffffffu, v.re, d), __shfl_up_sync(0xffffffffu, v.im, d)}; }
__device__ __forceinline__ float shfl_idx(float v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ double shfl_idx(double v, int l) { return __shfl_sync(0xffffffffu, v, l); }
__device__ __forceinline__ cplx shfl_idx(cplx v, int l) { return cplx{__shfl_sync(0xffffffffu, v.re, l), __shfl_sync(0xffffffffu, v.im, l)}; }

// 4 consecutive elements from a 16-byte aligned shared address
__device__ __forceinline__ void lds4(const int* p, int (&v)[4]) {
    const int4 t = *reinterpret_cast<const int4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void lds4(const long long* p, long long (&v)[4]) {
    const longlong2 a = reinterpret_cast<const longlong2*>(p)[0], b = reinterpret_cast<const longlong2*>(p)[1];
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
__device__ __forceinline__ void lds4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void lds4(const double* p, double (&v)[4]) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
__device__ __forceinline__ void lds4(const cplx* p, cplx (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double2 t = reinterpret_cast<const double2*>(p)[k];
        v[k] = cplx{t.x, t.y};
    }
}

// x gathers with an L2 evict-last hint: x is the one array the multiply re-reads (every column index points into it);
// the matrix stream passes through L2 evict-first, so x may stay resident when it fits (80 MB at config 4).
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float ld_x_keep(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_x_keep(const double* p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ cplx ld_x_keep(const cplx* p, uint64_t pol) {
    cplx v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.re), "=d"(v.im) : "l"(p), "l"(pol));
    return v;
}
template <bool GHOST, bool KEEP, class T, class Ti>
__device__ __forceinline__ T flat_x(const XView<T>& xv, Ti c, uint64_t pol) {
    const T* p = xv.own;
    if (GHOST) p = ((unsigned long long)((i64)c - xv.own_lo) < xv.own_n) ? xv.own : xv.gat;
    if (KEEP) return ld_x_keep(p + (i64)c, pol);
    return ld_x(p + (i64)c);
}

template <class T> struct FlatCfg { static constexpr int AHEAD = 4, CTAS = 4; };   // rounds of gathers issued before the first product
template <> struct FlatCfg<float> { static constexpr int AHEAD = 4, CTAS = 5; };
template <> struct FlatCfg<cplx> { static constexpr int AHEAD = 1, CTAS = 2; };

template <class T, class Ti, bool GHOST, bool KEEP>
__global__ void __launch_bounds__(FLAT_THREADS, FlatCfg<T>::CTAS) spmv_flat_kernel(const FlatArgs<T, Ti> a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned* sbits = reinterpret_cast<unsigned*>(smem_raw + 16);
    Ti* scol = reinterpret_cast<Ti*>(smem_raw + 16 + FLAT_WORDS * 4);
    T* sval = reinterpret_cast<T*>(scol + FLAT_CHUNK);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const i64 chunk = blockIdx.x;
    const i64 k0 = chunk * FLAT_CHUNK;
    const i64 left = a.nnz - k0;
    const int n_have = (int)(left < FLAT_CHUNK ? left : FLAT_CHUNK);  // > 0
    const int n_bulk = n_have & ~3;
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        const uint64_t pol = l2_evict_first_policy();
        mbar_expect_tx(bar, (uint32_t)(FLAT_WORDS * 4) + (uint32_t)n_bulk * (uint32_t)(sizeof(Ti) + sizeof(T)));
        bulk_g2s(sbits, a.bits + chunk * FLAT_WORDS, FLAT_WORDS * 4, bar, pol);
        if (n_bulk > 0) {
            bulk_g2s(scol, a.colval + k0, (uint32_t)n_bulk * (uint32_t)sizeof(Ti), bar, pol);
            bulk_g2s(sval, a.nzval + k0, (uint32_t)n_bulk * (uint32_t)sizeof(T), bar, pol);
        }
    }
    for (int k = n_bulk + tid; k < n_have; k += FLAT_THREADS) {  // the <= 3 entries a 16-byte copy cannot fetch (end of the arrays)
        scol[k] = a.colval[k0 + k];
        sval[k] = a.nzval[k0 + k];
    }
    __syncthreads();  // barrier initialised (and the tail written) before anyone waits
    mbar_wait(bar, 0);

    const int base = warp * FLAT_WCHUNK;
    if (base >= n_have) return;  // (no CTA-wide synchronisation below)
    const i64 wc = chunk * (FLAT_THREADS / 32) + warp;
    uint64_t xpol = 0;
    if (KEEP) xpol = l2_evict_last_policy();
    constexpr int AHEAD = FlatCfg<T>::AHEAD;
    const unsigned lt = (1u << lane) - 1u;
    const int first_flag = (int)(sbits[warp * (FLAT_WCHUNK / 32)] & 1u);
    i64 ord_open = __ldg(a.wrow + wc) - first_flag;  // ordinal of the row that is open in front of this round's first entry
    T carry = el_zero(T());                          // its partial sum so far (within this warp chunk)
    bool chunk_seen = false;                         // a row start has been met in this warp chunk
    auto row_of = [&](i64 ord) -> i64 { return a.row_map ? __ldg(a.row_map + ord) : ord; };

#pragma unroll
    for (int jb = 0; jb < FLAT_ROUNDS; jb += AHEAD) {
        T xg[AHEAD][4];
#pragma unroll
        for (int jj = 0; jj < AHEAD; ++jj) {  // all gathers of AHEAD rounds go out before the first product
            const int e = base + 128 * (jb + jj) + 4 * lane;
            Ti c[4];
            lds4(scol + e, c);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const Ti col = (e + k < n_have) ? c[k] : (Ti)a.safe_col;
                xg[jj][k] = flat_x<GHOST, KEEP, T, Ti>(a.xv, col, xpol);
            }
        }
#pragma unroll
        for (int jj = 0; jj < AHEAD; ++jj) {
            const int j = jb + jj;
            const int e = base + 128 * j + 4 * lane;
            T v[4], p[4];
            lds4(sval + e, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = (e + k < n_have) ? el_mul(v[k], xg[jj][k]) : el_zero(T());
            const unsigned f = (sbits[warp * (FLAT_WCHUNK / 32) + 4 * j + (lane >> 3)] >> (4 * (lane & 7))) & 0xFu;
            // row starts in front of my entries in this round, and in the whole round
            const unsigned b0 = __ballot_sync(0xffffffffu, f & 1u), b1 = __ballot_sync(0xffffffffu, f & 2u);
            const unsigned b2 = __ballot_sync(0xffffffffu, f & 4u), b3 = __ballot_sync(0xffffffffu, f & 8u);
            const int nbefore = __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
            const int ntotal = __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            // my 4 entries: head = what belongs to the row open in front of me, tail = what my last row start has so far;
            // rows that start and end inside my 4 entries are complete
            T acc = el_zero(T()), head = el_zero(T());
            bool seen = false;
            i64 cur = ord_open + nbefore;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if ((f >> k) & 1u) {
                    if (!seen) head = acc, seen = true;
                    else st_y(a.y + row_of(cur), acc);
                    cur += 1;
                    acc = p[k];
                } else {
                    acc = el_add(acc, p[k]);
                }
            }
            if (!seen) head = acc;
            // segmented inclusive scan of the tails over the lanes (a lane with a row start begins a new segment)
            T sv = acc;
            bool fl = seen;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const T vu = shfl_up(sv, d);
                const bool fu = __shfl_up_sync(0xffffffffu, (int)fl, d) != 0;
                if (lane >= d) {
                    if (!fl) sv = el_add(vu, sv);
                    fl = fl || fu;
                }
            }
            const T pv = shfl_up(sv, 1);
            const bool pf = __shfl_up_sync(0xffffffffu, (int)fl, 1) != 0;
            T cin = carry;
            bool before_seen = chunk_seen;
            if (lane > 0) {
                cin = pf ? pv : el_add(carry, pv);
                before_seen = chunk_seen || pf;
            }
            if (seen) {  // the row open in front of me ends at my first row start
                const T total = el_add(cin, head);
                if (before_seen) st_y(a.y + row_of(ord_open + nbefore), total);
                else a.heads[wc] = total;  // it began before this warp chunk (0 if the chunk begins with a row start)
            }
            const T lastv = shfl_idx(sv, 31);
            const bool lastf = __shfl_sync(0xffffffffu, (int)fl, 31) != 0;
            carry = lastf ? lastv : el_add(carry, lastv);
            chunk_seen = chunk_seen || lastf;
            ord_open += ntotal;
        }
    }
    if (lane == 0) {  // the row still open at the end of the warp chunk
        if (chunk_seen) st_y(a.y + row_of(ord_open), carry);  // it began here: its first part; the following heads are added by the fix-up
        else a.heads[wc] = carry;                             // the whole warp chunk lies inside one row
    }
}

// y[row] += heads of the warp chunks that continue `row`, in chunk order.  One thread per warp chunk; the thread of the
// FIRST continuation chunk of a row does the row.  max_run bounds the walk: longer rows belong to the long-row kernels,
// which overwrite y[row] afterwards.
template <class T>
__global__ void __launch_bounds__(256) flat_fixup_kernel(const unsigned* __restrict__ bits, const i64* __restrict__ wrow, const i64* __restrict__ row_map,
                                                         const T* __restrict__ heads, i64 n_wchunks, T* __restrict__ y, int max_run) {
    const i64 wc = (i64)blockIdx.x * 256 + threadIdx.x;
    if (wc >= n_wchunks || wc == 0) return;
    auto cont = [&](i64 c) { return (bits[c * (FLAT_WCHUNK / 32)] & 1u) == 0u; };
    if (!cont(wc)) return;
    const i64 ord = wrow[wc];
    if (cont(wc - 1) && wrow[wc - 1] == ord) return;  // an earlier chunk is the first continuation of this row
    const i64 row = row_map ? row_map[ord] : ord;
    T acc = y[row];
    int n = 0;
    for (i64 c = wc; c < n_wchunks && n < max_run; ++c, ++n) {
        if (c > wc && !(cont(c) && wrow[c] == ord)) break;
        acc = el_add(acc, heads[c]);
    }
    y[row] = acc;
}

template <class T>
__global__ void zero_rows_kernel(const i64* __restrict__ rows, i64 n, T* __restrict__ y) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[rows[i]] = el_zero(T());
}

// ------------------------------------------------------------------------------------------------------------------
// set-up (once per matrix)
// ------------------------------------------------------------------------------------------------------------------
template <class Ti>
__global__ void flat_bits_kernel(const Ti* __restrict__ rowptr, i64 nrows, unsigned* bits, unsigned char* nonempty) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const i64 b = (i64)rowptr[r] - 1, e = (i64)rowptr[r + 1] - 1;
    const bool ne = e > b;
    if (nonempty) nonempty[r] = ne ? 1 : 0;
    if (ne) atomicOr(bits + (b >> 5), 1u << (b & 31));
}

// wrow[wc] = the row holding stored entry wc * FLAT_WCHUNK: the last row whose first entry is at or before it (ties are
// empty rows followed by the non-empty one: the last of them); as an ordinal among the non-empty rows when ord_of_row is given
template <class Ti>
__global__ void flat_wrow_kernel(const Ti* __restrict__ rowptr, i64 nrows, i64 nnz, const i64* __restrict__ ord_of_row, i64* __restrict__ wrow, i64 n_wchunks) {
    const i64 wc = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (wc >= n_wchunks) return;
    i64 k = wc * FLAT_WCHUNK;
    if (k >= nnz) k = nnz - 1;
    i64 lo = 0, hi = nrows;  // first r with rowptr[r] - 1 > k
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if ((i64)rowptr[mid] - 1 > k) hi = mid;
        else lo = mid + 1;
    }
    const i64 r = lo - 1;
    wrow[wc] = ord_of_row ? ord_of_row[r] : r;
}

__global__ void flat_rowmap_kernel(const unsigned char* __restrict__ nonempty, const i64* __restrict__ ord_of_row, i64 nrows, i64 n_nonempty,
                                   i64* __restrict__ row_map, i64* __restrict__ empty_rows) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    if (nonempty[r]) row_map[ord_of_row[r]] = r;
    else empty_rows[r - ord_of_row[r]] = r;  // (rows before r that are non-empty = ord_of_row[r]; the rest are empty)
    (void)n_nonempty;
}

struct NonEmptyAsI64 {
    const unsigned char* p;
    __host__ __device__ i64 operator()(i64 i) const { return (i64)p[i]; }
};

cudaError_t flat_build(int itype, const void* rowptr, i64 nrows, i64 nnz, FlatData* F, cudaStream_t st) {
    *F = FlatData{};
    if (nnz <= 0 || nrows <= 0) return cudaSuccess;
    F->n_chunks = (nnz + FLAT_CHUNK - 1) / FLAT_CHUNK;
    F->n_wchunks = (nnz + FLAT_WCHUNK - 1) / FLAT_WCHUNK;  // warp chunks that hold entries
    cudaError_t e;
    if ((e = cudaMalloc(&F->d_bits, sizeof(unsigned) * (size_t)F->n_chunks * FLAT_WORDS)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(F->d_bits, 0, sizeof(unsigned) * (size_t)F->n_chunks * FLAT_WORDS, st)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&F->d_wrow, sizeof(i64) * (size_t)F->n_chunks * (FLAT_THREADS / 32))) != cudaSuccess) return e;
    unsigned char* d_ne = nullptr;
    if ((e = cudaMalloc(&d_ne, (size_t)nrows)) != cudaSuccess) return e;
    const int blocks = (int)((nrows + 255) / 256);
    if (itype == HPCLA_I32) flat_bits_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, nrows, F->d_bits, d_ne);
    else flat_bits_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, nrows, F->d_bits, d_ne);
    // ordinals among the non-empty rows (exclusive scan of the flags); only kept when some row is empty
    i64* d_ord = nullptr;
    if ((e = cudaMalloc(&d_ord, sizeof(i64) * (size_t)(nrows + 1))) != cudaSuccess) return e;
    {
        cub::TransformInputIterator<i64, NonEmptyAsI64, cub::CountingInputIterator<i64>> in(cub::CountingInputIterator<i64>(0), NonEmptyAsI64{d_ne});
        void* tmp = nullptr;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, d_ord, nrows, st);
        if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return e;
        e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, d_ord, nrows, st);
        if (e != cudaSuccess) return e;
        i64 last_ord = 0;
        unsigned char last_ne = 0;
        if ((e = cudaMemcpyAsync(&last_ord, d_ord + (nrows - 1), sizeof(i64), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(&last_ne, d_ne + (nrows - 1), 1, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        cudaFree(tmp);
        F->n_nonempty = last_ord + (last_ne ? 1 : 0);
    }
    const bool has_empty = F->n_nonempty < nrows;
    const int wblocks = (int)((F->n_wchunks + 255) / 256);
    if (itype == HPCLA_I32) flat_wrow_kernel<int><<<wblocks, 256, 0, st>>>((const int*)rowptr, nrows, nnz, has_empty ? d_ord : nullptr, F->d_wrow, F->n_wchunks);
    else flat_wrow_kernel<long long><<<wblocks, 256, 0, st>>>((const long long*)rowptr, nrows, nnz, has_empty ? d_ord : nullptr, F->d_wrow, F->n_wchunks);
    if (has_empty) {
        F->n_empty = nrows - F->n_nonempty;
        if ((e = cudaMalloc(&F->d_row_map, sizeof(i64) * (size_t)std::max<i64>(F->n_nonempty, 1))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&F->d_empty_rows, sizeof(i64) * (size_t)F->n_empty)) != cudaSuccess) return e;
        flat_rowmap_kernel<<<blocks, 256, 0, st>>>(d_ne, d_ord, nrows, F->n_nonempty, F->d_row_map, F->d_empty_rows);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    cudaFree(d_ne);
    cudaFree(d_ord);
    return cudaSuccess;
}

void flat_free(FlatData* F) {
    cudaFree(F->d_bits);
    cudaFree(F->d_wrow);
    cudaFree(F->d_row_map);
    cudaFree(F->d_empty_rows);
    cudaFree(F->d_heads);
    *F = FlatData{};
}

size_t flat_smem_bytes(int dtype, int itype) {
    const size_t ts = dtype == HPCLA_F32 ? 4 : dtype == HPCLA_F64 ? 8 : 16, is = itype == HPCLA_I32 ? 4 : 8;
    return 16 + (size_t)FLAT_WORDS * 4 + (size_t)FLAT_CHUNK * (is + ts);
}
int flat_max_run(i64 long_threshold) { return (int)(long_threshold / FLAT_WCHUNK) + 3; }

template <class T, class Ti, bool GHOST, bool KEEP>
static cudaError_t flat_launch_one(const FlatLaunch& L, const FlatArgs<T, Ti>& a, cudaStream_t st) {
    const size_t smem = flat_smem_bytes(L.dtype, L.itype);
    cudaError_t e;
    if ((e = ensure_smem<spmv_flat_kernel<T, Ti, GHOST, KEEP>>(smem, false)) != cudaSuccess) return e;
    spmv_flat_kernel<T, Ti, GHOST, KEEP><<<(unsigned)L.flat->n_chunks, FLAT_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <class T, class Ti>
static cudaError_t flat_typed(const FlatLaunch& L, cudaStream_t st) {
    const FlatData& F = *L.flat;
    if (F.n_chunks == 0) return cudaSuccess;
    FlatArgs<T, Ti> a;
    a.colval = (const Ti*)L.colval;
    a.nzval = (const T*)L.nzval;
    a.bits = F.d_bits;
    a.wrow = F.d_wrow;
    a.row_map = F.d_row_map;
    a.xv.own = L.x_own ? (const T*)L.x_own - L.own_lo : nullptr;
    a.xv.gat = L.gathered ? (const T*)L.gathered - 1 : nullptr;
    a.xv.own_lo = L.own_lo;
    a.xv.own_n = (unsigned long long)L.own_n;
    a.y = (T*)L.y;
    a.heads = (T*)F.d_heads;
    a.nnz = L.nnz;
    a.safe_col = L.own_n > 0 ? L.own_lo : 1;
    cudaError_t e;
    if (F.n_empty > 0) {
        zero_rows_kernel<T><<<(unsigned)((F.n_empty + 255) / 256), 256, 0, st>>>(F.d_empty_rows, F.n_empty, (T*)L.y);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (L.has_ghost) e = L.keep_x ? flat_launch_one<T, Ti, true, true>(L, a, st) : flat_launch_one<T, Ti, true, false>(L, a, st);
    else e = L.keep_x ? flat_launch_one<T, Ti, false, true>(L, a, st) : flat_launch_one<T, Ti, false, false>(L, a, st);
    if (e != cudaSuccess) return e;
    flat_fixup_kernel<T><<<(unsigned)((F.n_wchunks + 255) / 256), 256, 0, st>>>(F.d_bits, F.d_wrow, F.d_row_map, (const T*)F.d_heads, F.n_wchunks, (T*)L.y,
                                                                                flat_max_run(L.long_threshold));
    return cudaGetLastError();
}
cudaError_t launch_spmv_flat(const FlatLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(flat_typed, L, L, st); }

}  // namespace hpcla
