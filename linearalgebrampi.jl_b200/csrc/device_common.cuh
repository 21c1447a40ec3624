// device_common.cuh — device helpers shared by the multiply kernels (kernels.cu: SpMV, spmm.cu: sparse x dense):
// element arithmetic, streaming loads, the x view, the TMA staging of a tile and the row walk.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "device.h"

// A/B knobs of the row walk (alternate builds, tools/tune_spmv.py): predicated batches, gathers in flight per lane,
// resident CTAs per SM the ComplexF64 instantiations are register-allocated for
// (measured, profiles/r1e_tune_walk_variants.txt: predicated batches of 8 win on every stencil workload)
#ifndef HPCLA_WALK_PRED
#define HPCLA_WALK_PRED 1
#endif
#ifndef HPCLA_WALK_BATCH
#define HPCLA_WALK_BATCH 8
#endif
#ifndef HPCLA_CPLX_CTAS
#define HPCLA_CPLX_CTAS 3
#endif

namespace hpcla {

// ------------------------------------------------------------------------------------------------------------------
// element-type helpers
// ------------------------------------------------------------------------------------------------------------------
struct cplx {
    double re, im;
};

__device__ __forceinline__ float el_zero(float) { return 0.f; }
__device__ __forceinline__ double el_zero(double) { return 0.0; }
__device__ __forceinline__ cplx el_zero(cplx) { return cplx{0.0, 0.0}; }
// products are rounded on their own (__fmul_rn/__dmul_rn are never contracted into an FMA)
__device__ __forceinline__ float el_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double el_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ cplx el_mul(cplx a, cplx b) {
    return cplx{__dsub_rn(__dmul_rn(a.re, b.re), __dmul_rn(a.im, b.im)), __dadd_rn(__dmul_rn(a.re, b.im), __dmul_rn(a.im, b.re))};
}
__device__ __forceinline__ float el_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double el_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ cplx el_add(cplx a, cplx b) { return cplx{__dadd_rn(a.re, b.re), __dadd_rn(a.im, b.im)}; }

__device__ __forceinline__ float shfl_xor(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ cplx shfl_xor(cplx v, int m) {
    return cplx{__shfl_xor_sync(0xffffffffu, v.re, m), __shfl_xor_sync(0xffffffffu, v.im, m)};
}

// read-only-path scalar loads of x
__device__ __forceinline__ float ld_x(const float* p) { return __ldg(p); }
__device__ __forceinline__ double ld_x(const double* p) { return __ldg(p); }
__device__ __forceinline__ cplx ld_x(const cplx* p) {
    double2 t = __ldg(reinterpret_cast<const double2*>(p));
    return cplx{t.x, t.y};
}
// the same loads as volatile asm: the compiler may not sink them below a later barrier wait (plain ld.global.nc loads
// are "invariant" and do get moved across asm memory clobbers)
__device__ __forceinline__ float ld_x_pinned(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_x_pinned(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ cplx ld_x_pinned(const cplx* p) {
    cplx v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.re), "=d"(v.im) : "l"(p));
    return v;
}
// streaming (evict-first) scalar loads of the matrix, for the unaligned / tail cases
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    return cplx{t.x, t.y};
}
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ long long ld_stream(const long long* p) { return __ldcs(p); }

// 4 consecutive elements from a 16-byte aligned address, 128-bit streaming loads
__device__ __forceinline__ void ld4_stream(const float* p, float (&v)[4]) {
    float4 t = __ldcs(reinterpret_cast<const float4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void ld4_stream(const double* p, double (&v)[4]) {
    double2 a = __ldcs(reinterpret_cast<const double2*>(p));
    double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
__device__ __forceinline__ void ld4_stream(const cplx* p, cplx (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double2 t = __ldcs(reinterpret_cast<const double2*>(p) + k);
        v[k] = cplx{t.x, t.y};
    }
}
__device__ __forceinline__ void ld4_stream(const int* p, int (&v)[4]) {
    int4 t = __ldcs(reinterpret_cast<const int4*>(p));
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
}
__device__ __forceinline__ void ld4_stream(const long long* p, long long (&v)[4]) {
    longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(p));
    longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}
// 4 consecutive products to a 16-byte aligned shared address
__device__ __forceinline__ void st4_shared(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4_shared(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void st4_shared(cplx* p, const cplx (&v)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) reinterpret_cast<double2*>(p)[k] = make_double2(v[k].re, v[k].im);
}
__device__ __forceinline__ void st_y(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_y(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_y(cplx* p, cplx v) { __stcs(reinterpret_cast<double2*>(p), make_double2(v.re, v.im)); }

// general kernel, per element type: CTA size and 4-nonzero groups per lane and round
template <class T> struct TileCfg;
#ifndef HPCLA_GENERAL_F32_GROUPS
#define HPCLA_GENERAL_F32_GROUPS 2  // A/B knob
#endif
#ifndef HPCLA_GENERAL_MIN_CTAS
#define HPCLA_GENERAL_MIN_CTAS 1  // A/B knob: resident CTAs per SM the general kernel's registers are allocated for
#endif
template <> struct TileCfg<float> { static constexpr int THREADS = 256, GROUPS = HPCLA_GENERAL_F32_GROUPS; };
template <> struct TileCfg<double> { static constexpr int THREADS = 256, GROUPS = 2; };
template <> struct TileCfg<cplx> { static constexpr int THREADS = 256, GROUPS = 1; };
// row-walk kernel: CTA size, and the resident CTAs per SM the register allocation aims at
constexpr int ROW_THREADS = 256;
template <class T> struct RowCfg { static constexpr int CTAS = 8; };   // 32 registers per thread
#ifndef HPCLA_F64_CTAS
#define HPCLA_F64_CTAS 7  // A/B knob
#endif
template <> struct RowCfg<double> { static constexpr int CTAS = HPCLA_F64_CTAS; };  // 7: 36 registers
template <> struct RowCfg<cplx> { static constexpr int CTAS = HPCLA_CPLX_CTAS; };  // 48 (16-byte values)

// x addressing: own columns are read straight from x.v (no local copy into `gathered`), ghosts from `gathered`.
template <class T>
struct XView {
    const T* own;  // own[c] valid for own_lo <= c < own_lo + own_n   (pointer pre-shifted by -own_lo)
    const T* gat;  // gat[c] == gathered[c-1]                           (pointer pre-shifted by -1)
    i64 own_lo;
    unsigned long long own_n;
};
template <bool GHOST, class T, class Ti>
__device__ __forceinline__ T x_at(const XView<T>& xv, Ti c) {
    if (GHOST) {
        const T* p = ((unsigned long long)((i64)c - xv.own_lo) < xv.own_n) ? xv.own : xv.gat;
        return ld_x(p + (i64)c);
    }
    return ld_x(xv.own + (i64)c);
}

// A launch over an ascending tile list, described by value as at most 8 runs of consecutive tiles: the CTA's tile, hence
// its window of the nonzero stream, is then arithmetic on kernel parameters (no load in front of the first copy).
struct TileRuns {
    int n;         // runs in use (<= 8); 0: no such description, the tile records say everything
    int cta0[8];   // run j starts at CTA cta0[j] (ascending; unused runs: INT_MAX, so they never match)
    int skip[8];   // tile - CTA index within run j (= first tile of the run - cta0[j])
};
__device__ __forceinline__ i64 tile_of_cta(const TileRuns& runs) {
    // Constant indices only after unrolling (a run-time index would put a by-value copy of the arrays on the local-memory
    // stack), and no load depends on a comparison: the 16 parameter loads go out together, then 7 selects.
    const int cta = (int)blockIdx.x;
    int skip = runs.skip[0];
    if (runs.n > 1) {  // uniform: a single run (the common case) needs one parameter
#pragma unroll
        for (int j = 1; j < 8; ++j) skip = (cta >= runs.cta0[j]) ? runs.skip[j] : skip;
    }
    return (i64)(cta + skip);
}
static inline TileRuns launch_runs(int n, const int* cta0, const int* tile0) {
    TileRuns r;
    r.n = n;
    for (int j = 0; j < 8; ++j) {
        r.cta0[j] = j < n ? cta0[j] : 0x7fffffff;
        r.skip[j] = j < n ? tile0[j] - cta0[j] : 0;
    }
    return r;
}

template <class T, class Ti>
struct TileArgs {
    const Ti* rowptr;
    const Ti* colval;
    const T* nzval;
    XView<T> xv;
    T* y;
    const TileRec* recs;  // one record per CTA of this launch
    TileRuns runs;        // n > 0: the launch as runs of consecutive tiles (window arithmetic)
    int window;           // stored entries per tile window
    i64 nnz_total;
    i64 long_threshold;
    i64 safe_col;  // a column that is always valid to read (padding lanes)
    // fused dot(x, A x) of the row-walk kernel (CG's p.q): dot_x[r] pairs with row r; one partial per CTA (double)
    const T* dot_x;
    double* dot_out;
};

// ------------------------------------------------------------------------------------------------------------------
// Row-walk kernel: TMA-staged tiles.
// One elected thread issues three bulk asynchronous copies (cp.async.bulk, the 1-D TMA path: global -> shared through
// L2, bypassing the LSU/L1 wavefront pipeline that bounds the general kernel, evict-first in L2) for the tile's slice
// of colval, nzval and rowptr, completing on an mbarrier.  Then G lanes per row walk the staged row: with one lane per
// row, consecutive lanes handle consecutive ROWS, so for banded / stencil matrices the x gathers of one warp
// instruction hit consecutive addresses (coalesced), and the row is summed left to right exactly like the reference.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

template <class T>
__device__ __forceinline__ T block_sum(T v, T* sh /* [32] */) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = el_add(v, shfl_xor(v, m));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    T r = el_zero(T());
    if (warp == 0) {
        r = lane < (int)(blockDim.x >> 5) ? sh[lane] : el_zero(T());
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) r = el_add(r, shfl_xor(r, m));
    }
    return r;  // valid in warp 0
}

// Row walk: G lanes per row (interleaved: lane g takes entries g, g+G, ...), operands read from shared memory, x
// gathered per entry.  Lanes of a warp own consecutive rows, so on banded matrices the gathers of one warp instruction
// fall into a few contiguous runs (coalesced).  G = 1 sums left to right: the reference's order, bit for bit.
__device__ __forceinline__ double dot_term(float x, float y) { return (double)x * (double)y; }
__device__ __forceinline__ double dot_term(double x, double y) { return x * y; }
__device__ __forceinline__ double dot_term(cplx, cplx) { return 0.0; }  // the fused dot is for real types (CG)

// Returns this thread's share of sum_r dot_x[r] * y[r] over the rows it stored (0 when dot_x is null).
// DOT is a template parameter: the plain multiply must not carry the dot's live registers through its gather loop.
template <class T, class Ti, int THREADS, int G, bool GHOST, bool DOT = false>
__device__ __forceinline__ double rows_walk(const Ti* scol, const T* sval, const Ti* rp, i64 rp_off, const XView<T>& xv, T* __restrict__ y, i64 r0,
                                            i64 r1, i64 s4, int tid, const T* __restrict__ dot_x = nullptr) {
    constexpr int RPP = THREADS / G;
    const int lane = tid % G;
    double dot = 0.0;
    for (i64 base = r0; base < r1; base += RPP) {
        const i64 r = base + tid / G;
        const bool valid = r < r1;
        T acc = el_zero(T());
        if (valid) {
            const int b = (int)((i64)rp[r - rp_off] - 1 - s4);
            const int e = (int)((i64)rp[r - rp_off + 1] - 1 - s4);
            int k = b + lane;
            constexpr int B = HPCLA_WALK_BATCH;
#if HPCLA_WALK_PRED
            for (; k < e; k += B * G) {  // B predicated gathers in flight per lane, then the adds in order
                T p[B];
#pragma unroll
                for (int u = 0; u < B; ++u) {
                    const int kk = k + u * G;
                    p[u] = (kk < e) ? el_mul(sval[kk], x_at<GHOST, T, Ti>(xv, scol[kk])) : el_zero(T());
                }
#pragma unroll
                for (int u = 0; u < B; ++u)
                    if (k + u * G < e) acc = el_add(acc, p[u]);
            }
#else
            for (; k + (B - 1) * G < e; k += B * G) {  // B independent gathers in flight per lane, then the adds in order
                T p[B];
#pragma unroll
                for (int u = 0; u < B; ++u) p[u] = el_mul(sval[k + u * G], x_at<GHOST, T, Ti>(xv, scol[k + u * G]));
#pragma unroll
                for (int u = 0; u < B; ++u) acc = el_add(acc, p[u]);
            }
            for (; k < e; k += G) acc = el_add(acc, el_mul(sval[k], x_at<GHOST, T, Ti>(xv, scol[k])));
#endif
        }
        if (G > 1) {
#pragma unroll
            for (int m = G / 2; m >= 1; m >>= 1) acc = el_add(acc, shfl_xor(acc, m));
        }
        if (valid && lane == 0) {
            st_y(y + r, acc);
            if (DOT) dot += dot_term(ld_x(dot_x + r), acc);
        }
    }
    return dot;
}

// The matrix side of a tile, shared by the multiply kernels: what the bulk copies need (StageArgs), what they leave in
// shared memory (Staged).
template <class T, class Ti>
struct StageArgs {
    const Ti* rowptr;
    const Ti* colval;
    const T* nzval;
    const TileRec* recs;  // one record per CTA of this launch
    TileRuns runs;        // n > 0: the launch as runs of consecutive tiles (window arithmetic)
    int window;           // stored entries per tile window
    i64 nnz_total;
};
template <class T, class Ti>
struct Staged {
    const Ti* srp;   // row pointers of rows rp0 .. r1
    const Ti* scol;  // column indices, entry k <-> nonzero s4 + k
    const T* sval;
    i64 r0, r1, s4, rp0;
};

// One CTA per ROW-WALK tile (classified at set-up: every row of the tile is completely staged, its row pointers fit,
// and the lanes are well used).  Nothing but: the bulk copies, one wait, the walk.
template <class T, class Ti>
__device__ __forceinline__ Staged<T, Ti> stage_tile(const StageArgs<T, Ti>& a, int cap, int rp_cap, i64 rowptr_len, unsigned char* smem_raw) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    Ti* srp = reinterpret_cast<Ti*>(smem_raw + 16);
    Ti* scol = srp + rp_cap;
    T* sval = reinterpret_cast<T*>(scol + cap);
    const int tid = threadIdx.x;
    // A launch described as runs of consecutive tiles knows where its window of the nonzero stream starts without reading anything:
    // the bulk copies of the window are issued first, the tile record (rows, exact end) is fetched while they fly.
    const bool contig = a.runs.n > 0;
    i64 w0 = 0;
    int n_main = 0;
    if (contig) {
        w0 = tile_of_cta(a.runs) * (i64)a.window;
        const i64 left = (a.nnz_total - w0) & ~(i64)3;
        n_main = (int)(left < (i64)a.window ? (left > 0 ? left : 0) : (i64)a.window);
    }
    uint64_t pol = 0;
    if (tid == 0) {
        mbar_init(bar, contig ? 2 : 1);
        mbar_fence_init();
        pol = l2_evict_first_policy();
        if (contig) {
            mbar_expect_tx(bar, (uint32_t)n_main * (uint32_t)(sizeof(Ti) + sizeof(T)));
            if (n_main > 0) {
                bulk_g2s(scol, a.colval + w0, (uint32_t)n_main * (uint32_t)sizeof(Ti), bar, pol);
                bulk_g2s(sval, a.nzval + w0, (uint32_t)n_main * (uint32_t)sizeof(T), bar, pol);
            }
        }
    }
    // one 32-byte record per CTA (lists never hold tiles without rows)
    const longlong2 d0 = __ldg(reinterpret_cast<const longlong2*>(a.recs + blockIdx.x));
    const longlong2 d1 = __ldg(reinterpret_cast<const longlong2*>(a.recs + blockIdx.x) + 1);
    const i64 r0 = d0.x, r1 = d0.y;
    const i64 s = d1.x, e = d1.y;                // 0-based nonzero range of the tile's rows
    const i64 s4 = contig ? w0 : (s & ~(i64)3);  // staged from here: 16-byte aligned for every element width
    const int n_st = (int)(e - s4);              // <= cap (classification)
    const i64 avail = a.nnz_total - s4;
    const int n_bulk = (int)((((i64)n_st + 3) & ~(i64)3) <= avail ? (((i64)n_st + 3) & ~(i64)3) : (avail & ~(i64)3));
    const int n_rest = n_bulk - n_main;  // >= 0: a window ends inside the tile's last row (or at the end of the matrix)
    const i64 rp0 = r0 & ~(i64)3;
    const int rp_need = (int)(r1 + 1 - rp0);  // entries rp0 .. r1, <= rp_cap (classification)
    const i64 rp_avail = rowptr_len - rp0;
    const int rp_bulk = (int)((((i64)rp_need + 3) & ~(i64)3) <= rp_avail ? (((i64)rp_need + 3) & ~(i64)3) : (rp_avail & ~(i64)3));
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)n_rest * (uint32_t)(sizeof(Ti) + sizeof(T)) + (uint32_t)rp_bulk * (uint32_t)sizeof(Ti));
        if (n_rest > 0) {
            bulk_g2s(scol + n_main, a.colval + s4 + n_main, (uint32_t)n_rest * (uint32_t)sizeof(Ti), bar, pol);
            bulk_g2s(sval + n_main, a.nzval + s4 + n_main, (uint32_t)n_rest * (uint32_t)sizeof(T), bar, pol);
        }
        if (rp_bulk > 0) bulk_g2s(srp, a.rowptr + rp0, (uint32_t)rp_bulk * (uint32_t)sizeof(Ti), bar, pol);
    }
    // tails: the <= 3 elements a 16-byte copy would read past the end of an array (last tile of the matrix only)
    for (int k = n_bulk + tid; k < n_st; k += (int)blockDim.x) {
        scol[k] = a.colval[s4 + k];
        sval[k] = a.nzval[s4 + k];
    }
    for (int k = rp_bulk + tid; k < rp_need; k += (int)blockDim.x) srp[k] = a.rowptr[rp0 + k];
    __syncthreads();  // barrier initialised (and the tails written) before anyone waits
    mbar_wait(bar, 0);
    return Staged<T, Ti>{srp, scol, sval, r0, r1, s4, rp0};
}

template <class T, class Ti>
__device__ __forceinline__ StageArgs<T, Ti> stage_args(const TileArgs<T, Ti>& a) {
    return StageArgs<T, Ti>{a.rowptr, a.colval, a.nzval, a.recs, a.runs, a.window, a.nnz_total};
}

// ------------------------------------------------------------------------------------------------------------------
// host-side helpers of the launchers
// ------------------------------------------------------------------------------------------------------------------
// Opt-in dynamic shared memory per kernel instantiation and device (function attributes are per device).  The limit
// only ever grows: rank-threads of one process share the attribute and may ask for different sizes.
// carveout: prefer the maximum shared-memory carveout (row walk: x is gathered with coalesced accesses and needs
// little L1); the general kernel keeps the driver's default split, its scattered gathers live on L1.
template <auto Kernel>
static cudaError_t ensure_smem(size_t smem, bool carveout) {
    static size_t configured[64] = {};
    static bool carved[64] = {};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t& have = configured[dev & 63];
    if (have < smem && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        have = smem;
    }
    if (carveout && !carved[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        carved[dev & 63] = true;
    }
    return cudaSuccess;
}

#define HPCLA_DISPATCH(FN, L, ...)                                                                     \
    do {                                                                                               \
        if ((L).dtype == HPCLA_F32 && (L).itype == HPCLA_I32) return FN<float, int>(__VA_ARGS__);       \
        if ((L).dtype == HPCLA_F32 && (L).itype == HPCLA_I64) return FN<float, long long>(__VA_ARGS__); \
        if ((L).dtype == HPCLA_F64 && (L).itype == HPCLA_I32) return FN<double, int>(__VA_ARGS__);      \
        if ((L).dtype == HPCLA_F64 && (L).itype == HPCLA_I64) return FN<double, long long>(__VA_ARGS__);\
        if ((L).dtype == HPCLA_C128 && (L).itype == HPCLA_I32) return FN<cplx, int>(__VA_ARGS__);       \
        if ((L).dtype == HPCLA_C128 && (L).itype == HPCLA_I64) return FN<cplx, long long>(__VA_ARGS__); \
        return cudaErrorInvalidValue;                                                                  \
    } while (0)

}  // namespace hpcla
