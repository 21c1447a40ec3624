// transpose.cu — device-side distributed transpose: HPCSparseMatrix(transpose(A)) without leaving the GPU.
//
// Replaces TransposePlan(A) + execute_plan!(plan, A) (src/sparse.jl:1551-1744, 1756-1829): the reference copies the
// matrix to the host, buckets every nonzero by the owner of its column, exchanges (row, col) pairs (tag 10) and values
// (tag 11) through host buffers, and tuple-sorts all received entries (:1655) — minutes at BASELINE config 3.  Here:
//   1. expand_kernel: every local nonzero becomes (key, value) with key = owner-local row of A^T (its global column
//      minus the owner's first column) in the high 32 bits and the global column of A^T (its global row) in the low 32,
//      and is dropped straight into its owner's send range (per-owner cursors; order is irrelevant before the sort);
//   2. the per-owner counts travel as one grouped ncclSend/ncclRecv of 8-byte values, then the keys and the values as
//      two more groups, device to device;
//   3. one radix sort of the received keys (CUB DeviceRadixSort over the bits in use — library code, like cuBLAS for a
//      plain GEMM: this is the one-time set-up in front of the hot path, not the hot path) with the arrival index as
//      the payload orders the entries by (row, column): exactly the reference's tuple sort, entries are unique;
//   4. rowptr by one binary search per row, col_indices = sorted unique low words (a second sort + unique), colval by
//      one binary search per entry, nzval by one gather.
// Results are bit-identical to the host builder (hpcla_transpose_begin/finish) and to the oracle.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "device.h"

using namespace hpcla;

#define CU_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return fail(HPCLA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)


struct hpcla_dtb {
    int dtype = 0, itype = 0;
    i64 nrows = 0, nnz = 0, ncc = 0;
    void *d_rowptr = nullptr, *d_colval = nullptr, *d_nzval = nullptr;
    i64* d_col_indices = nullptr;
};

namespace {

typedef unsigned long long u64;

// owner of 1-based global column j in a partition of 1-based starts: last r with part[r] <= j, clamped (src/sparse.jl:1566-1579)
__device__ __forceinline__ int owner_of(const i64* __restrict__ part, int nranks, i64 j) {
    int lo = 0, hi = nranks;  // invariant: part[lo] <= j (part[0] = 1)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (part[mid] <= j) lo = mid;
        else hi = mid;
    }
    return lo;
}

// per compressed column: owner rank and key high word (row of A^T local to its owner)
__global__ void column_owner_kernel(const i64* __restrict__ col_indices, i64 ncc, const i64* __restrict__ col_partition, int nranks, int* __restrict__ owner,
                                    unsigned* __restrict__ local_row) {
    const i64 c = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncc) return;
    const i64 j = col_indices[c];
    const int o = owner_of(col_partition, nranks, j);
    owner[c] = o;
    local_row[c] = (unsigned)(j - col_partition[o]);
}

template <class Ti>
__global__ void __launch_bounds__(256) count_kernel(const Ti* __restrict__ colval, i64 nnz, const int* __restrict__ owner, int nranks, u64* __restrict__ counts) {
    __shared__ unsigned sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    for (i64 k = (i64)blockIdx.x * 256 + threadIdx.x; k < nnz; k += (i64)gridDim.x * 256) atomicAdd(&sh[owner[(i64)colval[k] - 1]], 1u);
    __syncthreads();
    if ((int)threadIdx.x < nranks && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (u64)sh[threadIdx.x]);
}

// one warp per row: entry -> (key, value) in its owner's send range
template <class T, class Ti>
__global__ void __launch_bounds__(256) expand_kernel(const Ti* __restrict__ rowptr, const Ti* __restrict__ colval, const T* __restrict__ nzval, i64 nrows,
                                                     i64 first_global_row /* 1-based */, const int* __restrict__ owner, const unsigned* __restrict__ local_row,
                                                     u64* __restrict__ cursors /* [nranks], start offsets */, u64* __restrict__ keys, T* __restrict__ vals) {
    const i64 r = (i64)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= nrows) return;
    const i64 b = (i64)rowptr[r] - 1, e = (i64)rowptr[r + 1] - 1;
    const u64 low = (u64)(first_global_row + r - 1);  // 0-based global row of A = column of A^T
    for (i64 k = b + lane; k < e; k += 32) {
        const i64 c = (i64)colval[k] - 1;
        const u64 slot = atomicAdd(&cursors[owner[c]], 1ull);
        keys[slot] = ((u64)local_row[c] << 32) | low;
        vals[slot] = nzval[k];
    }
}

__global__ void iota_kernel(unsigned* p, i64 n) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) p[k] = (unsigned)k;
}

template <class Ti>
__global__ void rowptr_kernel(const u64* __restrict__ keys, i64 nnz, i64 nrows, Ti* __restrict__ rowptr) {
    const i64 r = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nrows) return;
    const u64 target = (u64)r << 32;  // first key of row r
    i64 lo = 0, hi = nnz;
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (keys[mid] < target) lo = mid + 1;
        else hi = mid;
    }
    rowptr[r] = (Ti)(lo + 1);
}

__global__ void low_words_kernel(const u64* __restrict__ keys, i64 n, unsigned* __restrict__ out) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (unsigned)(keys[k] & 0xffffffffull);
}

template <class T, class Ti>
__global__ void finish_kernel(const u64* __restrict__ keys, const unsigned* __restrict__ perm, const T* __restrict__ vals_in, i64 nnz,
                              const unsigned* __restrict__ uniq, i64 ncc, Ti* __restrict__ colval, T* __restrict__ nzval) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const unsigned g = (unsigned)(keys[k] & 0xffffffffull);
    i64 lo = 0, hi = ncc;  // searchsortedfirst(col_indices, g)
    while (lo < hi) {
        const i64 mid = (lo + hi) >> 1;
        if (uniq[mid] < g) lo = mid + 1;
        else hi = mid;
    }
    colval[k] = (Ti)(lo + 1);
    nzval[k] = vals_in[perm[k]];
}

__global__ void widen_kernel(const unsigned* __restrict__ uniq, i64 n, i64* __restrict__ out) {
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = (i64)uniq[k] + 1;  // back to 1-based global indices
}

inline int blocks_of(i64 n, int t) { return (int)std::max<i64>(1, (n + t - 1) / t); }
inline int bits_for(u64 v) {
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

struct Scratch {  // frees everything on scope exit
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class P>
    cudaError_t alloc(P** p, size_t bytes) {
        cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(bytes, 16));
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

template <class T, class Ti>
int transpose_typed(hpcla_ctx* ctx, int rank, int nranks, bool has_comm, const i64* row_partition, const i64* col_partition, i64 nrows, i64 ncc, i64 nnz,
                    const Ti* d_rowptr, const Ti* d_colval, const i64* h_col_indices, const T* d_nzval, hpcla_dtb* out, cudaStream_t st) {
    Scratch S;
    if (nranks > 64) return fail(HPCLA_ERR_ARG, "hpcla_transpose_device: at most 64 ranks");
    if (row_partition[nranks] - 1 > 0xffffffffll || col_partition[nranks] - 1 > 0xffffffffll) return fail(HPCLA_ERR_ARG, "hpcla_transpose_device: global dimensions must fit 32 bits");
    // --- 1. owners and keys ---------------------------------------------------------------------------------------
    i64 *d_ci = nullptr, *d_cpart = nullptr;
    int* d_owner = nullptr;
    unsigned* d_lrow = nullptr;
    u64 *d_counts = nullptr, *d_cursors = nullptr;
    CU_TRY(S.alloc(&d_ci, sizeof(i64) * (size_t)ncc));
    CU_TRY(S.alloc(&d_cpart, sizeof(i64) * (size_t)(nranks + 1)));
    CU_TRY(S.alloc(&d_owner, sizeof(int) * (size_t)ncc));
    CU_TRY(S.alloc(&d_lrow, sizeof(unsigned) * (size_t)ncc));
    CU_TRY(S.alloc(&d_counts, sizeof(u64) * 64));
    CU_TRY(S.alloc(&d_cursors, sizeof(u64) * 64));
    if (ncc > 0) CU_TRY(cudaMemcpyAsync(d_ci, h_col_indices, sizeof(i64) * (size_t)ncc, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(d_cpart, col_partition, sizeof(i64) * (size_t)(nranks + 1), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(d_counts, 0, sizeof(u64) * 64, st));
    if (ncc > 0) column_owner_kernel<<<blocks_of(ncc, 256), 256, 0, st>>>(d_ci, ncc, d_cpart, nranks, d_owner, d_lrow);
    if (nnz > 0) count_kernel<Ti><<<(int)std::min<i64>(148 * 8, blocks_of(nnz, 256)), 256, 0, st>>>(d_colval, nnz, d_owner, nranks, d_counts);
    CU_TRY(cudaGetLastError());
    std::vector<u64> send_counts(64, 0);
    CU_TRY(cudaMemcpyAsync(send_counts.data(), d_counts, sizeof(u64) * 64, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    std::vector<i64> send_off((size_t)nranks + 1, 0);
    for (int q = 0; q < nranks; ++q) send_off[(size_t)q + 1] = send_off[(size_t)q] + (i64)send_counts[(size_t)q];
    if (send_off[(size_t)nranks] != nnz) return fail(HPCLA_ERR_STATE, "hpcla_transpose_device: owner counts do not add up");
    std::vector<u64> cursors(64, 0);
    for (int q = 0; q < nranks; ++q) cursors[(size_t)q] = (u64)send_off[(size_t)q];
    CU_TRY(cudaMemcpyAsync(d_cursors, cursors.data(), sizeof(u64) * 64, cudaMemcpyHostToDevice, st));
    u64* d_skeys = nullptr;
    T* d_svals = nullptr;
    CU_TRY(S.alloc(&d_skeys, sizeof(u64) * (size_t)nnz));
    CU_TRY(S.alloc(&d_svals, sizeof(T) * (size_t)nnz));
    if (nrows > 0 && nnz > 0) expand_kernel<T, Ti><<<blocks_of(nrows, 8), 256, 0, st>>>(d_rowptr, d_colval, d_nzval, nrows, row_partition[rank], d_owner, d_lrow, d_cursors, d_skeys, d_svals);
    CU_TRY(cudaGetLastError());
    // --- 2. exchange ----------------------------------------------------------------------------------------------
    std::vector<i64> recv_counts((size_t)nranks, 0);
    u64* d_rkeys = d_skeys;
    T* d_rvals = d_svals;
    i64 total = nnz;
    if (nranks > 1) {
        if (!has_comm) return fail(HPCLA_ERR_STATE, "hpcla_transpose_device: a multi-rank transpose needs an NCCL world");
        // counts: 8 bytes to / from every rank
        u64* d_rc = nullptr;
        CU_TRY(S.alloc(&d_rc, sizeof(u64) * 64));
        std::vector<i64> off8((size_t)nranks), len8((size_t)nranks, 8);
        for (int q = 0; q < nranks; ++q) off8[(size_t)q] = 8 * q;
        int rc = ctx_exchange_bytes(ctx, d_counts, off8.data(), len8.data(), d_rc, off8.data(), len8.data(), st);
        if (rc) return rc;
        std::vector<u64> rcounts(64, 0);
        CU_TRY(cudaMemcpyAsync(rcounts.data(), d_rc, sizeof(u64) * (size_t)nranks, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        std::vector<i64> recv_off((size_t)nranks + 1, 0);
        for (int q = 0; q < nranks; ++q) recv_counts[(size_t)q] = (i64)rcounts[(size_t)q], recv_off[(size_t)q + 1] = recv_off[(size_t)q] + recv_counts[(size_t)q];
        total = recv_off[(size_t)nranks];
        // receive buffers: a rank that cannot allocate them must not simply return — the others would wait for it in the
        // exchanges below — so the ranks agree on the worst status first
        cudaError_t ae = S.alloc(&d_rkeys, sizeof(u64) * (size_t)total);
        if (ae == cudaSuccess) ae = S.alloc(&d_rvals, sizeof(T) * (size_t)total);
        const int local_bad = (ae != cudaSuccess || total >= (i64)INT32_MAX) ? 1 : 0;
        if (ae != cudaSuccess) cudaGetLastError();
        int any_bad = 0;
        rc = ctx_agree_max(ctx, local_bad, &any_bad, st);
        if (rc) return rc;
        if (any_bad)
            return fail(local_bad ? (ae != cudaSuccess ? HPCLA_ERR_NOMEM : HPCLA_ERR_ARG) : HPCLA_ERR_STATE,
                        "hpcla_transpose_device: %s", local_bad ? (ae != cudaSuccess ? "out of device memory for the received entries" : "more than 2^31 - 1 entries on one rank")
                                                                : "another rank could not hold its share of the transpose");
        std::vector<i64> so((size_t)nranks), sb((size_t)nranks), ro((size_t)nranks), rb((size_t)nranks);
        for (int pass = 0; pass < 2; ++pass) {
            const i64 w = pass == 0 ? (i64)sizeof(u64) : (i64)sizeof(T);
            for (int q = 0; q < nranks; ++q) {
                so[(size_t)q] = send_off[(size_t)q] * w, sb[(size_t)q] = (i64)send_counts[(size_t)q] * w;
                ro[(size_t)q] = recv_off[(size_t)q] * w, rb[(size_t)q] = recv_counts[(size_t)q] * w;
            }
            rc = ctx_exchange_bytes(ctx, pass == 0 ? (const void*)d_skeys : (const void*)d_svals, so.data(), sb.data(), pass == 0 ? (void*)d_rkeys : (void*)d_rvals,
                                          ro.data(), rb.data(), st);
            if (rc) return rc;
        }
    }
    if (total >= (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_transpose_device: more than 2^31 - 1 entries on one rank");
    // --- 3. sort by (row of A^T, column of A^T) ----------------------------------------------------------------------
    const i64 nrows_t = col_partition[rank + 1] - col_partition[rank];
    u64* d_keys2 = nullptr;
    unsigned *d_perm = nullptr, *d_perm2 = nullptr;
    CU_TRY(S.alloc(&d_keys2, sizeof(u64) * (size_t)total));
    CU_TRY(S.alloc(&d_perm, sizeof(unsigned) * (size_t)total));
    CU_TRY(S.alloc(&d_perm2, sizeof(unsigned) * (size_t)total));
    if (total > 0) iota_kernel<<<blocks_of(total, 256), 256, 0, st>>>(d_perm, total);
    const int end_bit = std::min(64, 32 + bits_for((u64)std::max<i64>(nrows_t, 1)));
    cub::DoubleBuffer<u64> kb(d_rkeys, d_keys2);
    cub::DoubleBuffer<unsigned> pb(d_perm, d_perm2);
    size_t tmp_bytes = 0;
    void* d_tmp = nullptr;
    if (total > 0) {
        CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, pb, (int)total, 0, end_bit, st));
        CU_TRY(S.alloc(&d_tmp, tmp_bytes));
        CU_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kb, pb, (int)total, 0, end_bit, st));
    }
    const u64* d_sorted = kb.Current();
    const unsigned* d_order = pb.Current();
    // --- 4. CSR of the owned rows of A^T -----------------------------------------------------------------------------
    out->nrows = nrows_t;
    out->nnz = total;
    CU_TRY(cudaMalloc(&out->d_rowptr, std::max<size_t>(sizeof(Ti) * (size_t)(nrows_t + 1), 16)));
    CU_TRY(cudaMalloc(&out->d_colval, std::max<size_t>(sizeof(Ti) * (size_t)total, 16)));
    CU_TRY(cudaMalloc(&out->d_nzval, std::max<size_t>(sizeof(T) * (size_t)total, 16)));
    rowptr_kernel<Ti><<<blocks_of(nrows_t + 1, 256), 256, 0, st>>>(d_sorted, total, nrows_t, (Ti*)out->d_rowptr);
    // col_indices = unique!(sort(global columns of A^T)) (src/sparse.jl:501 applied to the transposed block)
    unsigned *d_low = nullptr, *d_low2 = nullptr, *d_uniq = nullptr;
    i64* d_nuniq = nullptr;
    CU_TRY(S.alloc(&d_low, sizeof(unsigned) * (size_t)total));
    CU_TRY(S.alloc(&d_low2, sizeof(unsigned) * (size_t)total));
    CU_TRY(S.alloc(&d_uniq, sizeof(unsigned) * (size_t)total));
    CU_TRY(S.alloc(&d_nuniq, sizeof(i64)));
    i64 n_uniq = 0;
    if (total > 0) {
        low_words_kernel<<<blocks_of(total, 256), 256, 0, st>>>(d_sorted, total, d_low);
        cub::DoubleBuffer<unsigned> lb(d_low, d_low2);
        const int low_bits = bits_for((u64)std::max<i64>(row_partition[nranks] - 1, 1));
        size_t t2 = 0;
        void* d_tmp2 = nullptr;
        CU_TRY(cub::DeviceRadixSort::SortKeys(nullptr, t2, lb, (int)total, 0, low_bits, st));
        CU_TRY(S.alloc(&d_tmp2, t2));
        CU_TRY(cub::DeviceRadixSort::SortKeys(d_tmp2, t2, lb, (int)total, 0, low_bits, st));
        size_t t3 = 0;
        void* d_tmp3 = nullptr;
        CU_TRY(cub::DeviceSelect::Unique(nullptr, t3, lb.Current(), d_uniq, d_nuniq, (int)total, st));
        CU_TRY(S.alloc(&d_tmp3, t3));
        CU_TRY(cub::DeviceSelect::Unique(d_tmp3, t3, lb.Current(), d_uniq, d_nuniq, (int)total, st));
        CU_TRY(cudaMemcpyAsync(&n_uniq, d_nuniq, sizeof(i64), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        finish_kernel<T, Ti><<<blocks_of(total, 256), 256, 0, st>>>(d_sorted, d_order, d_rvals, total, d_uniq, n_uniq, (Ti*)out->d_colval, (T*)out->d_nzval);
    }
    out->ncc = n_uniq;
    CU_TRY(cudaMalloc(&out->d_col_indices, std::max<size_t>(sizeof(i64) * (size_t)n_uniq, 16)));
    if (n_uniq > 0) widen_kernel<<<blocks_of(n_uniq, 256), 256, 0, st>>>(d_uniq, n_uniq, out->d_col_indices);
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaStreamSynchronize(st));
    return HPCLA_OK;
}

}  // namespace

extern "C" int hpcla_transpose_device(hpcla_ctx* ctx, int dtype, int itype, const int64_t* row_partition, const int64_t* col_partition, int64_t nrows_local,
                                      int64_t ncc, int64_t nnz, const void* d_rowptr, const void* d_colval, const int64_t* h_col_indices, const void* d_nzval,
                                      void* stream, hpcla_dtb** out) {
    if (!ctx || !row_partition || !col_partition || !out || nrows_local < 0 || ncc < 0 || nnz < 0 || !d_rowptr) return fail(HPCLA_ERR_ARG, "hpcla_transpose_device: bad arguments");
    if (!dtype_size(dtype) || !itype_size(itype)) return fail(HPCLA_ERR_ARG, "hpcla_transpose_device: unknown dtype/itype");
    int device = 0, rank = 0, nranks = 1, has_comm = 0;
    int rc = ctx_rank_info(ctx, &device, &rank, &nranks, &has_comm);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    hpcla_dtb* t = new hpcla_dtb();
    t->dtype = dtype;
    t->itype = itype;
    cudaStream_t st = (cudaStream_t)stream;
#define HPCLA_TR(TT, TI) rc = transpose_typed<TT, TI>(ctx, rank, nranks, has_comm != 0, row_partition, col_partition, nrows_local, ncc, nnz, (const TI*)d_rowptr, (const TI*)d_colval, h_col_indices, (const TT*)d_nzval, t, st)
    if (dtype == HPCLA_F32 && itype == HPCLA_I32) HPCLA_TR(float, int);
    else if (dtype == HPCLA_F32) HPCLA_TR(float, long long);
    else if (dtype == HPCLA_F64 && itype == HPCLA_I32) HPCLA_TR(double, int);
    else if (dtype == HPCLA_F64) HPCLA_TR(double, long long);
    else if (itype == HPCLA_I32) HPCLA_TR(double2, int);
    else HPCLA_TR(double2, long long);
#undef HPCLA_TR
    if (rc) {
        hpcla_dtb_destroy(t);
        return rc;
    }
    *out = t;
    return HPCLA_OK;
}

extern "C" int hpcla_dtb_sizes(const hpcla_dtb* t, int64_t* nrows_out, int64_t* nnz_out, int64_t* ncc_out) {
    if (!t) return fail(HPCLA_ERR_ARG, "hpcla_dtb_sizes: null");
    if (nrows_out) *nrows_out = t->nrows;
    if (nnz_out) *nnz_out = t->nnz;
    if (ncc_out) *ncc_out = t->ncc;
    return HPCLA_OK;
}

extern "C" int hpcla_dtb_result(const hpcla_dtb* t, void* d_rowptr_out, void* d_colval_out, int64_t* h_col_indices_out, void* d_nzval_out, void* stream) {
    if (!t) return fail(HPCLA_ERR_ARG, "hpcla_dtb_result: null");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t is = itype_size(t->itype), ts = dtype_size(t->dtype);
    if (d_rowptr_out) CU_TRY(cudaMemcpyAsync(d_rowptr_out, t->d_rowptr, is * (size_t)(t->nrows + 1), cudaMemcpyDeviceToDevice, st));
    if (d_colval_out && t->nnz > 0) CU_TRY(cudaMemcpyAsync(d_colval_out, t->d_colval, is * (size_t)t->nnz, cudaMemcpyDeviceToDevice, st));
    if (d_nzval_out && t->nnz > 0) CU_TRY(cudaMemcpyAsync(d_nzval_out, t->d_nzval, ts * (size_t)t->nnz, cudaMemcpyDeviceToDevice, st));
    if (h_col_indices_out && t->ncc > 0) CU_TRY(cudaMemcpyAsync(h_col_indices_out, t->d_col_indices, sizeof(i64) * (size_t)t->ncc, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return HPCLA_OK;
}

extern "C" void hpcla_dtb_destroy(hpcla_dtb* t) {
    if (!t) return;
    cudaFree(t->d_rowptr);
    cudaFree(t->d_colval);
    cudaFree(t->d_nzval);
    cudaFree(t->d_col_indices);
    delete t;
}
