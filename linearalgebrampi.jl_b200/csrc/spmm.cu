// spmm.cu — sparse x dense (multi-vector) kernels of the hot path's first caller, A * B::HPCMatrix.
#include "device_common.cuh"

namespace hpcla {

// ------------------------------------------------------------------------------------------------------------------
// Sparse x dense (multi-vector): C[:, k0 .. k0+K) = A * B[:, k0 .. k0+K) — replaces the reference's loop of ncols
// SpMVs with a column extraction each (Base.:*(A::HPCSparseMatrix, B::HPCMatrix), src/sparse.jl:2391-2413).
// The tile is staged once and walked once for K columns: the matrix stream is read once per K right-hand sides.
// B's own rows are read in place from its column-major local block (ldb); ghost rows come from a compact ROW-major
// buffer (all columns of one ghost row adjacent: one NCCL message per peer lands in place, no unpack).
// ------------------------------------------------------------------------------------------------------------------
template <class T>
struct XViewM {
    const T* own;    // own[c + k * ldb] for own_lo <= c < own_lo + own_n   (pointer pre-shifted by -own_lo)
    const T* ghost;  // ghost[g * ncols + k], g = 0-based ghost number       (compact, row-major)
    i64 ldb;
    i64 own_lo;
    unsigned long long own_n;
    int ncols;  // row length of the ghost buffer
};
template <bool GHOST, int K, class T, class Ti>
__device__ __forceinline__ void xm_at(const XViewM<T>& xv, Ti c, int k0, T (&out)[K]) {
    const i64 ci = (i64)c;
    if (GHOST && (unsigned long long)(ci - xv.own_lo) >= xv.own_n) {
        const i64 g = ci < xv.own_lo ? ci - 1 : ci - 1 - (i64)xv.own_n;
        const T* p = xv.ghost + g * (i64)xv.ncols + k0;
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = ld_x(p + k);
    } else {
        const T* p = xv.own + ci + (i64)k0 * xv.ldb;
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = ld_x(p + (i64)k * xv.ldb);
    }
}

template <class T, class Ti>
struct TileArgsM {
    StageArgs<T, Ti> st;
    XViewM<T> xv;
    T* y;     // column-major, ldc
    i64 ldc;
    int k0;   // first column of this launch
};

// Registers: left to itself ptxas squeezes this kernel into 32 registers (8 CTAs per SM) and, depending on a handful of
// live values, pays for it by reusing one destination register for all gathers of a batch, i.e. by serialising them
// (measured: 495 -> 601 us per 4 columns).  The bound below gives it room for the whole batch in flight.
#ifndef HPCLA_SPMM_MIN_CTAS
#define HPCLA_SPMM_MIN_CTAS 6  // A/B knob
#endif
template <class T, class Ti, int G, bool GHOST, int K>
__global__ void __launch_bounds__(ROW_THREADS, K >= 8 ? 4 : HPCLA_SPMM_MIN_CTAS) spmm_rowwalk_kernel(const TileArgsM<T, Ti> a, int cap, int rp_cap, i64 rowptr_len) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Staged<T, Ti> st = stage_tile<T, Ti>(a.st, cap, rp_cap, rowptr_len, smem_raw);
    constexpr int RPP = ROW_THREADS / G;
#ifndef HPCLA_SPMM_EB4
#define HPCLA_SPMM_EB4 2  // A/B knob: entries per batch at K = 4
#endif
    constexpr int EB = K >= 8 ? 1 : K >= 4 ? HPCLA_SPMM_EB4 : 4;  // entries per batch: EB * K gathers in flight per lane
    const int tid = threadIdx.x, lane = tid % G;
    for (i64 base = st.r0; base < st.r1; base += RPP) {
        const i64 r = base + tid / G;
        const bool valid = r < st.r1;
        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = el_zero(T());
        if (valid) {
            const int b = (int)((i64)st.srp[r - st.rp0] - 1 - st.s4);
            const int e = (int)((i64)st.srp[r - st.rp0 + 1] - 1 - st.s4);
            for (int j = b + lane; j < e; j += EB * G) {
                T xg[EB][K], v[EB];
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const int jj = j + u * G;
                    const bool ok = jj < e;
                    v[u] = ok ? st.sval[jj] : el_zero(T());
                    xm_at<GHOST, K, T, Ti>(a.xv, ok ? st.scol[jj] : st.scol[j], a.k0, xg[u]);
                }
#pragma unroll
                for (int u = 0; u < EB; ++u)
                    if (j + u * G < e) {
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], el_mul(v[u], xg[u][k]));
                    }
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
            for (int k = 0; k < K; ++k) st_y(a.y + r + (i64)(a.k0 + k) * a.ldc, acc[k]);
        }
    }
}

// general tiles (and every long row): one warp per row straight from global memory, lanes stride the row, K columns
template <class T, class Ti, bool GHOST, int K>
__global__ void __launch_bounds__(256) spmm_rows_warp_kernel(const TileArgsM<T, Ti> a) {
    const longlong2 d0 = __ldg(reinterpret_cast<const longlong2*>(a.st.recs + blockIdx.x));
    const i64 r0 = d0.x, r1 = d0.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (i64 r = r0 + warp; r < r1; r += 8) {
        const i64 b = (i64)__ldg(a.st.rowptr + r) - 1, e = (i64)__ldg(a.st.rowptr + r + 1) - 1;
        T acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = el_zero(T());
        for (i64 j = b + lane; j < e; j += 32) {
            T xg[K];
            const T v = ld_stream(a.st.nzval + j);
            xm_at<GHOST, K, T, Ti>(a.xv, ld_stream(a.st.colval + j), a.k0, xg);
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], el_mul(v, xg[k]));
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = el_add(acc[k], shfl_xor(acc[k], m));
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) st_y(a.y + r + (i64)(a.k0 + k) * a.ldc, acc[k]);
        }
    }
}

// sendbuf[(i) * ncols + k] = B[idx[i] - 1 + k * ldb]: all peers, all columns in one launch (row-major messages)
template <class T>
__global__ void pack_rows_kernel(const T* __restrict__ B, i64 ldb, const i64* __restrict__ idx, i64 n, int ncols, T* __restrict__ out) {
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * ncols) return;
    const i64 i = t / ncols;
    const int k = (int)(t - i * ncols);
    out[t] = B[idx[i] - 1 + (i64)k * ldb];
}


// ==================================================================================================================
// host-side launchers
// ==================================================================================================================
template <class T, class Ti>
static TileArgsM<T, Ti> tile_args_m(const SpmmLaunch& L) {
    TileArgsM<T, Ti> a;
    a.st = StageArgs<T, Ti>{(const Ti*)L.rowptr, (const Ti*)L.colval, (const T*)L.nzval, L.recs, launch_runs(L.n_runs, L.run_cta0, L.run_tile0), L.shape.window, L.nnz};
    a.xv.own = L.b_own ? (const T*)L.b_own - L.own_lo : nullptr;
    a.xv.ghost = (const T*)L.ghost;
    a.xv.ldb = L.ldb;
    a.xv.own_lo = L.own_lo;
    a.xv.own_n = (unsigned long long)L.own_n;
    a.xv.ncols = L.ncols;
    a.y = (T*)L.c;
    a.ldc = L.ldc;
    a.k0 = L.k0;
    return a;
}

template <class T, class Ti, int G, int K>
static cudaError_t spmm_rowwalk_launch(const SpmmLaunch& L, const TileArgsM<T, Ti>& a, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (L.has_ghost) {
        if ((e = ensure_smem<spmm_rowwalk_kernel<T, Ti, G, true, K>>(smem, true)) != cudaSuccess) return e;
        spmm_rowwalk_kernel<T, Ti, G, true, K><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
    } else {
        if ((e = ensure_smem<spmm_rowwalk_kernel<T, Ti, G, false, K>>(smem, true)) != cudaSuccess) return e;
        spmm_rowwalk_kernel<T, Ti, G, false, K><<<L.n_launch, ROW_THREADS, smem, st>>>(a, L.shape.cap, L.shape.rp_cap, L.nrows + 1);
    }
    return cudaGetLastError();
}

template <class T, class Ti, int K>
static cudaError_t spmm_rowwalk_lanes(const SpmmLaunch& L, cudaStream_t st) {
    const TileArgsM<T, Ti> a = tile_args_m<T, Ti>(L);
    const size_t smem = rowwalk_smem_bytes(L.dtype, L.itype, L.shape);
    switch (L.shape.lanes) {
        case 1: return spmm_rowwalk_launch<T, Ti, 1, K>(L, a, smem, st);
        case 2: return spmm_rowwalk_launch<T, Ti, 2, K>(L, a, smem, st);
        case 4: return spmm_rowwalk_launch<T, Ti, 4, K>(L, a, smem, st);
        case 8: return spmm_rowwalk_launch<T, Ti, 8, K>(L, a, smem, st);
    }
    return cudaErrorInvalidValue;  // wider walks go through the warp-per-row kernel (spmm_supports_rowwalk)
}

template <class T, class Ti>
static cudaError_t spmm_rowwalk_typed(const SpmmLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    if (L.kn == 8) return spmm_rowwalk_lanes<T, Ti, 8>(L, st);
    if (L.kn == 4) return spmm_rowwalk_lanes<T, Ti, 4>(L, st);
    if (L.kn == 1) return spmm_rowwalk_lanes<T, Ti, 1>(L, st);
    return cudaErrorInvalidValue;
}

template <class T, class Ti>
static cudaError_t spmm_rows_typed(const SpmmLaunch& L, cudaStream_t st) {
    if (L.n_launch <= 0) return cudaSuccess;
    const TileArgsM<T, Ti> a = tile_args_m<T, Ti>(L);
    if (L.kn == 8) {
        if (L.has_ghost) spmm_rows_warp_kernel<T, Ti, true, 8><<<L.n_launch, 256, 0, st>>>(a);
        else spmm_rows_warp_kernel<T, Ti, false, 8><<<L.n_launch, 256, 0, st>>>(a);
    } else if (L.kn == 4) {
        if (L.has_ghost) spmm_rows_warp_kernel<T, Ti, true, 4><<<L.n_launch, 256, 0, st>>>(a);
        else spmm_rows_warp_kernel<T, Ti, false, 4><<<L.n_launch, 256, 0, st>>>(a);
    } else if (L.kn == 1) {
        if (L.has_ghost) spmm_rows_warp_kernel<T, Ti, true, 1><<<L.n_launch, 256, 0, st>>>(a);
        else spmm_rows_warp_kernel<T, Ti, false, 1><<<L.n_launch, 256, 0, st>>>(a);
    } else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

bool spmm_supports_rowwalk(const TileShape& shape) { return shape.lanes >= 1 && shape.lanes <= 8; }
cudaError_t launch_spmm_rowwalk(const SpmmLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(spmm_rowwalk_typed, L, L, st); }
cudaError_t launch_spmm_rows(const SpmmLaunch& L, cudaStream_t st) { HPCLA_DISPATCH(spmm_rows_typed, L, L, st); }

cudaError_t launch_pack_rows(int dtype, const void* B, i64 ldb, const i64* idx, i64 n, int ncols, void* out, cudaStream_t st) {
    if (n == 0 || ncols == 0) return cudaSuccess;
    const i64 total = n * ncols;
    const int b = (int)((total + 255) / 256);
    if (dtype == HPCLA_F32) pack_rows_kernel<float><<<b, 256, 0, st>>>((const float*)B, ldb, idx, n, ncols, (float*)out);
    else if (dtype == HPCLA_F64) pack_rows_kernel<double><<<b, 256, 0, st>>>((const double*)B, ldb, idx, n, ncols, (double*)out);
    else if (dtype == HPCLA_C128) pack_rows_kernel<double2><<<b, 256, 0, st>>>((const double2*)B, ldb, idx, n, ncols, (double2*)out);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace hpcla
