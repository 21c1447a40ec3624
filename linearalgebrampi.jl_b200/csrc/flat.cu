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

// Warps per CTA (one persistent CTA per SM).  What bounds this kernel is the number of scattered x gathers an SM keeps
// in flight, and those live in the L1 lines that the shared-memory carve-out leaves over (measured, tools/gather_probe.cu
// and profiles/r2a_*, r2b_*: 0.87 gathers per clock and SM with ~190 KB of L1, a third of that with 60 KB).  So the
// kernel stages as little as it can: every warp keeps a two-stage ring holding only the COLUMN INDICES and row-start flags
// of its 512-entry warp chunks (they are needed a whole chunk ahead, to issue the gathers); the values are fetched with
// 128-bit streaming loads straight into registers, one chunk ahead as well.  As many warps as fit in ~60 KB.
template <class T, class Ti> struct FlatCfg {
    static constexpr int STAGE = 64 + FLAT_WCHUNK * (int)sizeof(Ti);
    static constexpr int FIT = (60 * 1024) / (2 * STAGE);
    static constexpr int BY_REGS = sizeof(T) <= 4 ? 16 : sizeof(T) <= 8 ? 12 : 6;  // x and values of two chunks live in registers
#ifndef HPCLA_FLAT_MAX_WARPS
#define HPCLA_FLAT_MAX_WARPS 16  // A/B knob
#endif
    static constexpr int W0 = FIT < BY_REGS ? FIT : BY_REGS;
    static constexpr int WARPS = W0 < HPCLA_FLAT_MAX_WARPS ? W0 : HPCLA_FLAT_MAX_WARPS;
    static constexpr bool PREFETCH = sizeof(T) <= 8;  // gathers and values of the next chunk in flight while this one is reduced
};

// 4 consecutive values from a 16-byte aligned global address, streaming (each value is used once)
__device__ __forceinline__ void ldg4_stream(const float* p, float (&v)[4]) { ld4_stream(p, v); }
__device__ __forceinline__ void ldg4_stream(const double* p, double (&v)[4]) { ld4_stream(p, v); }
__device__ __forceinline__ void ldg4_stream(const cplx* p, cplx (&v)[4]) { ld4_stream(p, v); }

template <class T, class Ti, bool GHOST, bool KEEP>
__global__ void __launch_bounds__(FlatCfg<T, Ti>::WARPS * 32, 1) spmv_flat_kernel(const FlatArgs<T, Ti> a, i64 n_wchunks) {
    using Cfg = FlatCfg<T, Ti>;
    constexpr int WARPS = Cfg::WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + 2 * warp;  // this warp's two barriers
    unsigned char* ring = smem_raw + 256 + (size_t)warp * 2 * Cfg::STAGE;
    const i64 total_warps = (i64)gridDim.x * WARPS;
    i64 wc = (i64)blockIdx.x * WARPS + warp;  // this warp takes chunks wc, wc + total_warps, ...
    if (wc >= n_wchunks) return;              // (warps are independent: no CTA-wide synchronisation anywhere)
    uint64_t pol = 0, xpol = 0;
    if (lane == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
        pol = l2_evict_first_policy();
    }
    if (KEEP) xpol = l2_evict_last_policy();
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    auto stage_bits = [&](int s) { return reinterpret_cast<unsigned*>(ring + (size_t)s * Cfg::STAGE); };
    auto stage_col = [&](int s) { return reinterpret_cast<Ti*>(ring + (size_t)s * Cfg::STAGE + 64); };
    auto have = [&](i64 c) { const i64 left = a.nnz - c * FLAT_WCHUNK; return (int)(left < FLAT_WCHUNK ? left : FLAT_WCHUNK); };
    // fetch the row-start flags and column indices of warp chunk c into stage s: two bulk copies on the stage's barrier
    auto issue = [&](i64 c, int s) {
        const int n_have = have(c), n_bulk = n_have & ~3;
        const i64 k0 = c * FLAT_WCHUNK;
        if (lane == 0) {
            mbar_expect_tx(bars + s, 64u + (uint32_t)n_bulk * (uint32_t)sizeof(Ti));
            bulk_g2s(stage_bits(s), a.bits + c * (FLAT_WCHUNK / 32), 64, bars + s, pol);
            if (n_bulk > 0) bulk_g2s(stage_col(s), a.colval + k0, (uint32_t)n_bulk * (uint32_t)sizeof(Ti), bars + s, pol);
        }
        if (lane < n_have - n_bulk) stage_col(s)[n_bulk + lane] = a.colval[k0 + n_bulk + lane];  // (end of the arrays)
    };
    // all 16 gathers of a lane go out together, and the 16 values behind them
    auto gather = [&](i64 c, int s, T (&xg)[FLAT_ROUNDS][4], T (&vg)[FLAT_ROUNDS][4]) {
        const int n_have = have(c);
        const Ti* scol = stage_col(s);
        const T* gval = a.nzval + c * FLAT_WCHUNK;
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) {
            const int e = 128 * j + 4 * lane;
            Ti cc[4];
            lds4(scol + e, cc);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const Ti col = (e + k < n_have) ? cc[k] : (Ti)a.safe_col;
                xg[j][k] = flat_x<GHOST, KEEP, T, Ti>(a.xv, col, xpol);
            }
        }
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) {
            const int e = 128 * j + 4 * lane;
            if (e + 4 <= n_have) {
                ldg4_stream(gval + e, vg[j]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) vg[j][k] = (e + k < n_have) ? ld_stream(gval + e + k) : el_zero(T());
            }
        }
    };
    auto row_of = [&](i64 ord) -> i64 { return a.row_map ? __ldg(a.row_map + ord) : ord; };
    // reduce warp chunk c (flags staged in s, x and values in registers) by row: register-level segmented scan over 4 rounds of
    // 128 entries.  The segment boundaries of each scan come from one ballot (no flag shuffles), and the four rounds' scans
    // are independent of each other (only the short carry chain at the end is sequential), so they are issued interleaved:
    // five dependent shuffle levels per chunk instead of twenty.  The common case of at most one row start among a
    // lane's 4 entries is branch-free; rows shorter than 4 entries take a uniform slow path.
    auto reduce = [&](i64 c, int s, const T (&xg)[FLAT_ROUNDS][4], const T (&vg)[FLAT_ROUNDS][4], i64 wrow_c) {
        const int n_have = have(c);
        const unsigned* sbits = stage_bits(s);
        const int first_flag = (int)(sbits[0] & 1u);
        T head[FLAT_ROUNDS], sv[FLAT_ROUNDS];
        unsigned f[FLAT_ROUNDS], seen_mask[FLAT_ROUNDS];
        int nbefore[FLAT_ROUNDS], ntotal[FLAT_ROUNDS], seg_lo[FLAT_ROUNDS];
        bool any_short = false;
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) {
            const int e = 128 * j + 4 * lane;
            f[j] = (e < n_have) ? (sbits[4 * j + (lane >> 3)] >> (4 * (lane & 7))) & 0xFu : 0u;
            any_short = any_short || __popc(f[j]) > 1;
        }
        const bool slow = __any_sync(0xffffffffu, any_short);  // rows that start AND end inside one lane's 4 entries (rare)
        i64 ord0 = wrow_c - first_flag;  // ordinal of the row that is open in front of the chunk's first entry
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) {
            const int e = 128 * j + 4 * lane;
            T p[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = (e + k < n_have) ? el_mul(vg[j][k], xg[j][k]) : el_zero(T());
            // row starts in front of my entries in this round, and in the whole round
            const unsigned b0 = __ballot_sync(0xffffffffu, f[j] & 1u), b1 = __ballot_sync(0xffffffffu, f[j] & 2u);
            const unsigned b2 = __ballot_sync(0xffffffffu, f[j] & 4u), b3 = __ballot_sync(0xffffffffu, f[j] & 8u);
            nbefore[j] = __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt);
            ntotal[j] = __popc(b0) + __popc(b1) + __popc(b2) + __popc(b3);
            const bool seen = f[j] != 0u;
            // my 4 entries: head = what belongs to the row open in front of me, tail = what my last row start has so far
            T tail;
            if (slow) {
                T acc = el_zero(T());
                head[j] = el_zero(T());
                bool sn = false;
                i64 ord_open = ord0;
#pragma unroll
                for (int jj = 0; jj < FLAT_ROUNDS; ++jj)
                    if (jj < j) ord_open += ntotal[jj];
                i64 cur = ord_open + nbefore[j];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if ((f[j] >> k) & 1u) {
                        if (!sn) head[j] = acc, sn = true;
                        else st_y(a.y + row_of(cur), acc);  // a row inside my 4 entries: complete
                        cur += 1;
                        acc = p[k];
                    } else {
                        acc = el_add(acc, p[k]);
                    }
                }
                if (!sn) head[j] = acc;
                tail = acc;
            } else {
                const int pos = seen ? __ffs((int)f[j]) - 1 : 4;  // my (only) row start
                head[j] = el_zero(T());
                tail = el_zero(T());
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    head[j] = el_add(head[j], k < pos ? p[k] : el_zero(T()));
                    tail = el_add(tail, k >= pos ? p[k] : el_zero(T()));
                }
                if (!seen) tail = head[j];
            }
            seen_mask[j] = __ballot_sync(0xffffffffu, seen);
            const unsigned below = seen_mask[j] & (lt | (1u << lane));  // lanes with a row start, up to and including me
            seg_lo[j] = below ? 31 - __clz((int)below) : 0;            // first lane of my segment
            sv[j] = tail;
        }
        // segmented inclusive scans of the tails over the lanes, the four rounds interleaved
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T vu[FLAT_ROUNDS];
#pragma unroll
            for (int j = 0; j < FLAT_ROUNDS; ++j) vu[j] = shfl_up(sv[j], d);
#pragma unroll
            for (int j = 0; j < FLAT_ROUNDS; ++j)
                if (lane - d >= seg_lo[j]) sv[j] = el_add(vu[j], sv[j]);
        }
        T pv[FLAT_ROUNDS], lastv[FLAT_ROUNDS];
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) pv[j] = shfl_up(sv[j], 1), lastv[j] = shfl_idx(sv[j], 31);
        // the carry chain across the rounds
        i64 ord_open = ord0;
        T carry = el_zero(T());   // partial sum of the open row so far (within this warp chunk)
        bool chunk_seen = false;  // a row start has been met in this warp chunk
#pragma unroll
        for (int j = 0; j < FLAT_ROUNDS; ++j) {
            const bool seen = f[j] != 0u;
            const bool pf = (seen_mask[j] & lt) != 0u;
            T cin = carry;
            bool before_seen = chunk_seen;
            if (lane > 0) {
                cin = pf ? pv[j] : el_add(carry, pv[j]);
                before_seen = chunk_seen || pf;
            }
            if (seen) {  // the row open in front of me ends at my first row start
                const T total = el_add(cin, head[j]);
                if (before_seen) st_y(a.y + row_of(ord_open + nbefore[j]), total);
                else a.heads[c] = total;  // it began before this warp chunk (0 if the chunk begins with a row start)
            }
            const bool lastf = seen_mask[j] != 0u;
            carry = lastf ? lastv[j] : el_add(carry, lastv[j]);
            chunk_seen = chunk_seen || lastf;
            ord_open += ntotal[j];
        }
        if (lane == 0) {  // the row still open at the end of the warp chunk
            if (chunk_seen) st_y(a.y + row_of(ord_open), carry);  // it began here: its first part; the following heads are added by the fix-up
            else a.heads[c] = carry;                              // the whole warp chunk lies inside one row
        }
    };

    // software pipeline per warp: copy of chunk i+2 | gathers + values of chunk i+1 | reduction of chunk i
    uint32_t ph0 = 0, ph1 = 0;  // phase parity to wait for, per stage
    issue(wc, 0);
    if (wc + total_warps < n_wchunks) issue(wc + total_warps, 1);
    __syncwarp();
    T xa[FLAT_ROUNDS][4], va[FLAT_ROUNDS][4], xb[FLAT_ROUNDS][4], vb[FLAT_ROUNDS][4];
    i64 wrow_a = __ldg(a.wrow + wc), wrow_b = 0;
    mbar_wait(bars, ph0);
    ph0 ^= 1;
    if (Cfg::PREFETCH) gather(wc, 0, xa, va);
    int s = 0;
    for (; wc < n_wchunks; wc += total_warps) {
        __syncwarp();  // (the tail entries a lane may have written into a stage are visible to the others)
        const i64 nxt = wc + total_warps;
        const bool more = nxt < n_wchunks;
        if (more) {
            wrow_b = __ldg(a.wrow + nxt);
            if (s == 0) mbar_wait(bars + 1, ph1), ph1 ^= 1;
            else mbar_wait(bars, ph0), ph0 ^= 1;
            if (Cfg::PREFETCH) gather(nxt, s ^ 1, xb, vb);
        }
        if (!Cfg::PREFETCH) gather(wc, s, xa, va);
        reduce(wc, s, xa, va, wrow_a);
        __syncwarp();  // every lane is done with stage s before it is refilled
        if (nxt + total_warps < n_wchunks) issue(nxt + total_warps, s);
        if (Cfg::PREFETCH) {
#pragma unroll
            for (int j = 0; j < FLAT_ROUNDS; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) xa[j][k] = xb[j][k], va[j][k] = vb[j][k];
        }
        wrow_a = wrow_b;
        s ^= 1;
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
    struct Temps {  // scratch of this function, released on every way out (what lands in *F is released by flat_free)
        void* p[3] = {nullptr, nullptr, nullptr};
        ~Temps() {
            for (void* q : p) cudaFree(q);
        }
    } temps;
    if ((e = cudaMalloc(&F->d_bits, sizeof(unsigned) * (size_t)F->n_chunks * FLAT_WORDS)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(F->d_bits, 0, sizeof(unsigned) * (size_t)F->n_chunks * FLAT_WORDS, st)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&F->d_wrow, sizeof(i64) * (size_t)F->n_chunks * (FLAT_THREADS / 32))) != cudaSuccess) return e;
    unsigned char* d_ne = nullptr;
    if ((e = cudaMalloc(&d_ne, (size_t)nrows)) != cudaSuccess) return e;
    temps.p[0] = d_ne;
    const int blocks = (int)((nrows + 255) / 256);
    if (itype == HPCLA_I32) flat_bits_kernel<int><<<blocks, 256, 0, st>>>((const int*)rowptr, nrows, F->d_bits, d_ne);
    else flat_bits_kernel<long long><<<blocks, 256, 0, st>>>((const long long*)rowptr, nrows, F->d_bits, d_ne);
    // ordinals among the non-empty rows (exclusive scan of the flags); only kept when some row is empty
    i64* d_ord = nullptr;
    if ((e = cudaMalloc(&d_ord, sizeof(i64) * (size_t)(nrows + 1))) != cudaSuccess) return e;
    temps.p[1] = d_ord;
    {
        cub::TransformInputIterator<i64, NonEmptyAsI64, cub::CountingInputIterator<i64>> in(cub::CountingInputIterator<i64>(0), NonEmptyAsI64{d_ne});
        void* tmp = nullptr;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, d_ord, nrows, st);
        if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) return e;
        temps.p[2] = tmp;
        e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, d_ord, nrows, st);
        if (e != cudaSuccess) return e;
        i64 last_ord = 0;
        unsigned char last_ne = 0;
        if ((e = cudaMemcpyAsync(&last_ord, d_ord + (nrows - 1), sizeof(i64), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(&last_ne, d_ne + (nrows - 1), 1, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
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

template <class T, class Ti>
static size_t flat_smem() { return 256 + (size_t)FlatCfg<T, Ti>::WARPS * 2 * FlatCfg<T, Ti>::STAGE; }
int flat_max_run(i64 long_threshold) { return (int)(long_threshold / FLAT_WCHUNK) + 3; }

// opt-in dynamic shared memory and a carve-out that leaves the rest of the SM's 228 KB to L1 (the gathers' lines)
template <auto Kernel>
static cudaError_t flat_configure(size_t smem) {
    static bool done[64] = {};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (done[dev & 63]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int pct = (int)((smem + 1024) * 100 / (228 * 1024)) + 1;  // the driver rounds up to the next carve-out it supports
    e = cudaFuncSetAttribute(Kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return e;
    done[dev & 63] = true;
    return cudaSuccess;
}

template <class T, class Ti, bool GHOST, bool KEEP>
static cudaError_t flat_launch_one(const FlatLaunch& L, const FlatArgs<T, Ti>& a, cudaStream_t st) {
    const size_t smem = flat_smem<T, Ti>();
    cudaError_t e;
    if ((e = flat_configure<spmv_flat_kernel<T, Ti, GHOST, KEEP>>(smem)) != cudaSuccess) return e;
    static int sms[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!sms[dev & 63]) cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    constexpr int WARPS = FlatCfg<T, Ti>::WARPS;
    i64 grid = sms[dev & 63] > 0 ? sms[dev & 63] : 148;  // one persistent CTA per SM
    const i64 need = (L.flat->n_wchunks + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    spmv_flat_kernel<T, Ti, GHOST, KEEP><<<(unsigned)grid, WARPS * 32, smem, st>>>(a, L.flat->n_wchunks);
    return cudaGetLastError();
}

template <class T, class Ti>
static cudaError_t flat_typed(const FlatLaunch& L, cudaStream_t st) {
    const FlatData& F = *L.flat;
    if (F.n_wchunks == 0) return cudaSuccess;
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
