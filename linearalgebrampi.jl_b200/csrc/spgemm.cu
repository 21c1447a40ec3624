// spgemm.cu — sparse x sparse, A * B::HPCSparseMatrix: memoised symbolic product + numeric phase on the device.
//
// The reference (src/sparse.jl:991-1059) gathers the rows of B that A's columns reference (MatrixPlan, :579-922), copies
// everything to the host and calls SparseArrays' CSC product `plan.AT * A_csc` on every call: the symbolic work is redone
// each time and the values make two PCIe trips.  Here the structure of C and, for every stored entry of C, the list of
// (entry of A, entry of the gathered B) pairs that contribute to it are computed ONCE per pair of structures (host,
// threads); a product with new values is then one kernel:
//     C.nzval[d] = sum over the pairs t of d, in ascending order of the shared index k, of Bg.nzval[ib[t]] * A.nzval[ia[t]]
// — the order and the operand order (b * a) of SparseArrays' `spmatmul` applied to (B^T)(A^T), so the values match the
// reference bit for bit; structural entries that cancel to zero are kept, as there.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "device.h"

using namespace hpcla;

#define CU_TRY(expr)                                                                                         \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess) return fail(HPCLA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct hpcla_spgemm {
    i64 nrows = 0, nnz = 0, ncc = 0, nterms = 0;
    std::vector<i64> rowptr;       // [nrows+1], 1-based
    std::vector<i64> cols_global;  // [nnz], 1-based global columns, ascending within a row
    std::vector<i64> col_indices;  // [ncc] sorted unique global columns
    std::vector<i64> term_ptr;     // [nnz+1], 0-based offsets into ia / ib
    std::vector<int> ia, ib;       // [nterms] 0-based entries of A.nzval / of the gathered B values
    // device copies (first numeric call)
    int device = -1;
    i64* d_term_ptr = nullptr;
    int *d_ia = nullptr, *d_ib = nullptr;
};

namespace {

struct Triple {
    i64 c;
    int ia, ib;
};

template <class Ti>
int symbolic_typed(i64 nrows, const Ti* a_rowptr, const Ti* a_colval, i64 n_rows_g, const i64* bg_rowptr, const i64* bg_cols, hpcla_spgemm* P) {
    for (i64 i = 0; i < nrows; ++i)
        for (i64 j = (i64)a_rowptr[i] - 1; j < (i64)a_rowptr[i + 1] - 1; ++j)
            if ((i64)a_colval[j] < 1 || (i64)a_colval[j] > n_rows_g) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_symbolic: A refers to gathered row %lld of %lld", (long long)a_colval[j], (long long)n_rows_g);
    const i64 nnz_a = nrows > 0 ? (i64)a_rowptr[nrows] - 1 : 0;
    const i64 nnz_b = bg_rowptr[n_rows_g] - 1;
    if (nnz_a >= (i64)INT32_MAX || nnz_b >= (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_symbolic: operands with 2^31 or more stored entries are not supported");
    P->nrows = nrows;
    P->rowptr.assign((size_t)nrows + 1, 1);
    // One pass, one sort per row: thread t takes a contiguous block of rows and appends to its own buffers (output
    // columns, terms per output entry, the term lists); the blocks are then laid end to end.
    const int nt = host_threads(nnz_a * 8);
    struct Part {
        std::vector<i64> cols, tcount;  // per output entry: global column, number of terms
        std::vector<int> ia, ib;
    };
    std::vector<Part> parts((size_t)nt);
    auto work = [&](int t) {
        Part& W = parts[(size_t)t];
        const i64 lo = nrows * t / nt, hi = nrows * (t + 1) / nt;
        std::vector<Triple> tr;
        for (i64 i = lo; i < hi; ++i) {
            tr.clear();
            for (i64 j = (i64)a_rowptr[i] - 1; j < (i64)a_rowptr[i + 1] - 1; ++j) {
                const i64 g = (i64)a_colval[j] - 1;
                for (i64 q = bg_rowptr[g] - 1; q < bg_rowptr[g + 1] - 1; ++q) tr.push_back(Triple{bg_cols[q], (int)j, (int)q});
            }
            // by output column, stable in k: A's entries ascend in k, so does the collection order
            std::stable_sort(tr.begin(), tr.end(), [](const Triple& x, const Triple& y) { return x.c < y.c; });
            i64 nout = 0;
            for (size_t u = 0; u < tr.size(); ++u) {
                if (u == 0 || tr[u].c != tr[u - 1].c) {
                    W.cols.push_back(tr[u].c);
                    W.tcount.push_back(0);
                    ++nout;
                }
                W.tcount.back() += 1;
                W.ia.push_back(tr[u].ia);
                W.ib.push_back(tr[u].ib);
            }
            P->rowptr[(size_t)i + 1] = nout;  // row counts for now
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (i64 i = 0; i < nrows; ++i) P->rowptr[(size_t)i + 1] += P->rowptr[(size_t)i];
    P->nnz = P->rowptr[(size_t)nrows] - 1;
    P->nterms = 0;
    for (const Part& W : parts) P->nterms += (i64)W.ia.size();
    if (P->nterms >= (i64)1 << 40) return fail(HPCLA_ERR_NOMEM, "hpcla_spgemm_symbolic: %lld product terms", (long long)P->nterms);
    P->cols_global.resize((size_t)P->nnz);
    P->term_ptr.assign((size_t)P->nnz + 1, 0);
    P->ia.resize((size_t)P->nterms);
    P->ib.resize((size_t)P->nterms);
    {
        i64 d = 0, tpos = 0;
        for (Part& W : parts) {
            std::copy(W.cols.begin(), W.cols.end(), P->cols_global.begin() + d);
            for (size_t u = 0; u < W.tcount.size(); ++u) {
                P->term_ptr[(size_t)d + u] = tpos;
                tpos += W.tcount[u];
            }
            std::copy(W.ia.begin(), W.ia.end(), P->ia.begin() + (tpos - (i64)W.ia.size()));
            std::copy(W.ib.begin(), W.ib.end(), P->ib.begin() + (tpos - (i64)W.ib.size()));
            d += (i64)W.cols.size();
            Part().cols.swap(W.cols), Part().tcount.swap(W.tcount), Part().ia.swap(W.ia), Part().ib.swap(W.ib);  // release as we go
        }
        P->term_ptr[(size_t)P->nnz] = tpos;
    }
    // col_indices = unique!(sort(global columns)) (src/sparse.jl:1023): a presence map over the columns in use
    if (P->nnz > 0) {
        i64 cmax = 0;
        for (i64 c : P->cols_global) cmax = std::max(cmax, c);
        std::vector<unsigned char> seen((size_t)cmax + 1, 0);
        for (i64 c : P->cols_global) seen[(size_t)c] = 1;
        for (i64 c = 1; c <= cmax; ++c)
            if (seen[(size_t)c]) P->col_indices.push_back(c);
    }
    P->ncc = (i64)P->col_indices.size();
    return HPCLA_OK;
}

template <class T>
__device__ __forceinline__ T mul_ba(T b, T a);
template <> __device__ __forceinline__ float mul_ba(float b, float a) { return __fmul_rn(b, a); }
template <> __device__ __forceinline__ double mul_ba(double b, double a) { return __dmul_rn(b, a); }
template <> __device__ __forceinline__ double2 mul_ba(double2 b, double2 a) {
    return make_double2(__dsub_rn(__dmul_rn(b.x, a.x), __dmul_rn(b.y, a.y)), __dadd_rn(__dmul_rn(b.x, a.y), __dmul_rn(b.y, a.x)));
}
__device__ __forceinline__ float add2(float x, float y) { return __fadd_rn(x, y); }
__device__ __forceinline__ double add2(double x, double y) { return __dadd_rn(x, y); }
__device__ __forceinline__ double2 add2(double2 x, double2 y) { return make_double2(__dadd_rn(x.x, y.x), __dadd_rn(x.y, y.y)); }
__device__ __forceinline__ float zero_of(float) { return 0.f; }
__device__ __forceinline__ double zero_of(double) { return 0.0; }
__device__ __forceinline__ double2 zero_of(double2) { return make_double2(0.0, 0.0); }

// one thread per stored entry of C: its terms are adjacent, neighbouring threads read neighbouring term ranges
template <class T>
__global__ void __launch_bounds__(256) spgemm_numeric_kernel(i64 nnz, const i64* __restrict__ term_ptr, const int* __restrict__ ia, const int* __restrict__ ib,
                                                             const T* __restrict__ a, const T* __restrict__ bg, T* __restrict__ c) {
    const i64 d = (i64)blockIdx.x * 256 + threadIdx.x;
    if (d >= nnz) return;
    T acc = zero_of(T());
    for (i64 t = term_ptr[d]; t < term_ptr[d + 1]; ++t) acc = add2(acc, mul_ba(bg[ib[t]], a[ia[t]]));
    c[d] = acc;
}

}  // namespace

extern "C" int hpcla_spgemm_symbolic(int itype, int64_t nrows_local, const void* a_rowptr, const void* a_colval, int64_t n_gathered_rows, const int64_t* bg_rowptr,
                                     const int64_t* bg_cols, hpcla_spgemm** out) {
    if (!out || nrows_local < 0 || n_gathered_rows < 0 || !a_rowptr || !bg_rowptr) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_symbolic: bad arguments");
    hpcla_spgemm* P = new hpcla_spgemm();
    int rc;
    if (itype == HPCLA_I32) rc = symbolic_typed<int32_t>(nrows_local, (const int32_t*)a_rowptr, (const int32_t*)a_colval, n_gathered_rows, bg_rowptr, bg_cols, P);
    else if (itype == HPCLA_I64) rc = symbolic_typed<int64_t>(nrows_local, (const int64_t*)a_rowptr, (const int64_t*)a_colval, n_gathered_rows, bg_rowptr, bg_cols, P);
    else rc = fail(HPCLA_ERR_ARG, "hpcla_spgemm_symbolic: unknown index type %d", itype);
    if (rc) {
        delete P;
        return rc;
    }
    *out = P;
    return HPCLA_OK;
}

extern "C" int hpcla_spgemm_sizes(const hpcla_spgemm* P, int64_t* nnz_out, int64_t* ncc_out, int64_t* nterms_out) {
    if (!P) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_sizes: null");
    if (nnz_out) *nnz_out = P->nnz;
    if (ncc_out) *ncc_out = P->ncc;
    if (nterms_out) *nterms_out = P->nterms;
    return HPCLA_OK;
}

extern "C" int hpcla_spgemm_structure(const hpcla_spgemm* P, int itype, void* rowptr_out, void* colval_out, int64_t* col_indices_out) {
    if (!P || !rowptr_out) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_structure: null");
    if (itype != HPCLA_I32 && itype != HPCLA_I64) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_structure: unknown index type %d", itype);
    if (itype == HPCLA_I32 && P->nnz >= (i64)INT32_MAX) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_structure: the product does not fit Int32 row pointers");
    for (i64 i = 0; i <= P->nrows; ++i) {
        if (itype == HPCLA_I32) ((int32_t*)rowptr_out)[i] = (int32_t)P->rowptr[(size_t)i];
        else ((int64_t*)rowptr_out)[i] = P->rowptr[(size_t)i];
    }
    // compress_map[global column] = position in col_indices (src/sparse.jl:1026-1034)
    std::vector<i64> compress_map(P->col_indices.empty() ? 1 : (size_t)P->col_indices.back() + 1, 0);
    for (size_t k = 0; k < P->col_indices.size(); ++k) compress_map[(size_t)P->col_indices[k]] = (i64)k + 1;
    for (i64 d = 0; d < P->nnz; ++d) {
        const i64 local = compress_map[(size_t)P->cols_global[(size_t)d]];
        if (itype == HPCLA_I32) ((int32_t*)colval_out)[d] = (int32_t)local;
        else ((int64_t*)colval_out)[d] = local;
    }
    if (col_indices_out && P->ncc > 0) std::memcpy(col_indices_out, P->col_indices.data(), sizeof(i64) * (size_t)P->ncc);
    return HPCLA_OK;
}

extern "C" int hpcla_spgemm_numeric(hpcla_spgemm* P, hpcla_ctx* ctx, int dtype, const void* d_a_nzval, const void* d_bg_nzval, void* d_c_nzval, void* stream) {
    if (!P || !ctx || !dtype_size(dtype)) return fail(HPCLA_ERR_ARG, "hpcla_spgemm_numeric: bad arguments");
    int device = 0, rank = 0, nranks = 1, has_comm = 0;
    int rc = ctx_rank_info(ctx, &device, &rank, &nranks, &has_comm);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (P->nnz == 0) return HPCLA_OK;
    if (!P->d_term_ptr) {  // the memoised term lists go to the device once
        P->device = device;
        CU_TRY(cudaMalloc(&P->d_term_ptr, sizeof(i64) * (size_t)(P->nnz + 1)));
        CU_TRY(cudaMalloc(&P->d_ia, sizeof(int) * (size_t)std::max<i64>(P->nterms, 1)));
        CU_TRY(cudaMalloc(&P->d_ib, sizeof(int) * (size_t)std::max<i64>(P->nterms, 1)));
        CU_TRY(cudaMemcpy(P->d_term_ptr, P->term_ptr.data(), sizeof(i64) * (size_t)(P->nnz + 1), cudaMemcpyHostToDevice));
        if (P->nterms > 0) {
            CU_TRY(cudaMemcpy(P->d_ia, P->ia.data(), sizeof(int) * (size_t)P->nterms, cudaMemcpyHostToDevice));
            CU_TRY(cudaMemcpy(P->d_ib, P->ib.data(), sizeof(int) * (size_t)P->nterms, cudaMemcpyHostToDevice));
        }
    } else if (P->device != device) {
        return fail(HPCLA_ERR_STATE, "hpcla_spgemm_numeric: the plan lives on device %d", P->device);
    }
    const int blocks = (int)((P->nnz + 255) / 256);
    if (dtype == HPCLA_F32) spgemm_numeric_kernel<float><<<blocks, 256, 0, st>>>(P->nnz, P->d_term_ptr, P->d_ia, P->d_ib, (const float*)d_a_nzval, (const float*)d_bg_nzval, (float*)d_c_nzval);
    else if (dtype == HPCLA_F64) spgemm_numeric_kernel<double><<<blocks, 256, 0, st>>>(P->nnz, P->d_term_ptr, P->d_ia, P->d_ib, (const double*)d_a_nzval, (const double*)d_bg_nzval, (double*)d_c_nzval);
    else spgemm_numeric_kernel<double2><<<blocks, 256, 0, st>>>(P->nnz, P->d_term_ptr, P->d_ia, P->d_ib, (const double2*)d_a_nzval, (const double2*)d_bg_nzval, (double2*)d_c_nzval);
    CU_TRY(cudaGetLastError());
    return HPCLA_OK;
}

extern "C" void hpcla_spgemm_destroy(hpcla_spgemm* P) {
    if (!P) return;
    if (P->device >= 0) {
        cudaSetDevice(P->device);
        cudaFree(P->d_term_ptr);
        cudaFree(P->d_ia);
        cudaFree(P->d_ib);
    }
    delete P;
}

// all-to-all of byte ranges between device buffers (the value exchange of a MatrixPlan: requested rows of B.nzval)
extern "C" int hpcla_exchange_bytes(hpcla_ctx* ctx, const void* d_send, const int64_t* send_off, const int64_t* send_bytes, void* d_recv, const int64_t* recv_off,
                                    const int64_t* recv_bytes, void* stream) {
    if (!ctx || !send_off || !send_bytes || !recv_off || !recv_bytes) return fail(HPCLA_ERR_ARG, "hpcla_exchange_bytes: null");
    int device = 0, rank = 0, nranks = 1, has_comm = 0;
    int rc = ctx_rank_info(ctx, &device, &rank, &nranks, &has_comm);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    return ctx_exchange_bytes(ctx, d_send, send_off, send_bytes, d_recv, recv_off, recv_bytes, (cudaStream_t)stream);
}
