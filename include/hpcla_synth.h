/* hpcla_synth.h — deterministic synthetic inputs shared by the tests, the benchmark and the CPU baseline
 * (SURVEY.md §8d).  Host-only helpers in a library of their own, hpcla_synth/libhpcla_synth.so (g++, no CUDA), so that
 * the CPU reference arm can draw the same inputs without mapping the product library; they are not part of the
 * interface the reference would bind (the reference builds its matrices with SparseArrays, e.g.
 * tools/benchmark_vs_petsc.jl:42-49 for the 2-D Laplacian).  Type codes: dtype 0 Float32, 1 Float64, 2 ComplexF64;
 * itype 0 Int32, 1 Int64 (as in hpcla_b200.h).  Every int-returning function returns 0 on success.
 *
 * Every generator emits the rows [row_begin, row_end) (0-based global rows) of the matrix as the input of
 * HPCSparseMatrix_local (src/sparse.jl:454): 1-based rowptr (Ti), 1-based GLOBAL columns ascending within a row (Ti),
 * values (T).  u(k) = (splitmix64(seed + k) >> 11) * 2^-53.
 */
#ifndef HPCLA_SYNTH_H
#define HPCLA_SYNTH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    HPCLA_SYNTH_LAPLACE2D_5PT = 0, /* nx x ny grid, g = ix + nx*iy, diag 4, neighbours -1, Dirichlet truncation (nz ignored) */
    HPCLA_SYNTH_POISSON3D_7PT = 1, /* nx x ny x nz grid, g = ix + nx*iy + nx*ny*iz, diag 6, neighbours -1                      */
    HPCLA_SYNTH_STENCIL3D_27PT = 2 /* centre 26, others -1, plus 0.1*((2u1-1) + i(2u2-1)) keyed on (g*27+d): A^T != A  */
};

int64_t hpcla_synth_stencil_rows(int kind, int64_t nx, int64_t ny, int64_t nz);
int64_t hpcla_synth_stencil_nnz(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end);
int hpcla_synth_stencil_fill(int kind, int64_t nx, int64_t ny, int64_t nz, int dtype, int itype, int64_t row_begin,
                             int64_t row_end, void* rowptr, void* global_cols, void* nzval);

/* power-law rows: L_g = min(max_len, n, floor(7 * (1 - u(g))^(-1/1.5))) (Pareto, alpha 2.5, k_min 7); the k-th column
 * of row g is floor((k + u(g,k)) * n / L_g) (stratified: ascending and distinct); values 2u - 1. */
int64_t hpcla_synth_powerlaw_nnz(int64_t n, uint64_t seed, int64_t max_len, int64_t row_begin, int64_t row_end);
int hpcla_synth_powerlaw_fill(int64_t n, uint64_t seed, int64_t max_len, int dtype, int itype, int64_t row_begin,
                              int64_t row_end, void* rowptr, void* global_cols, void* nzval);

/* x[g] = 2u(g) - 1 for g in [begin, end) (0-based); ComplexF64 adds an independent imaginary part. */
int hpcla_synth_vector(int dtype, uint64_t seed, int64_t begin, int64_t end, void* out);

/* message of the last failure on the calling thread */
const char* hpcla_synth_last_error(void);
/* worker threads the generators may use on the calling thread's behalf (0 = one per hardware thread, at most 32);
 * the CPU reference arm sets 1 inside each of its workers so that a worker first-touches its own arrays */
void hpcla_synth_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
