// hpcla_synth.cpp — deterministic synthetic matrices and vectors (see include/hpcla_synth.h).
// Its own small host-only library (libhpcla_synth.so, g++): the tests, bench.py and the CPU reference arm all draw
// their inputs from it, and the reference arm must not map the product library to do so.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../include/hpcla_synth.h"

typedef int64_t i64;
enum { HPCLA_OK = 0, HPCLA_ERR_ARG = 1 };                  // status values as in include/hpcla_b200.h
enum { HPCLA_F32 = 0, HPCLA_F64 = 1, HPCLA_C128 = 2, HPCLA_I32 = 0, HPCLA_I64 = 1 };  // type codes as in include/hpcla_b200.h

namespace {
thread_local std::string g_error;
thread_local int g_max_threads = 0;  // 0: one per hardware thread (at most 32); hpcla_synth_set_threads
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}
int host_threads(i64 work_items) {
    if (work_items < (i64)1 << 20) return 1;
    if (g_max_threads > 0) return g_max_threads;
    unsigned hc = std::thread::hardware_concurrency();
    int t = hc ? (int)hc : 4;
    return t > 32 ? 32 : t;
}

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double unit(uint64_t seed, uint64_t k) { return (double)(splitmix64(seed + k) >> 11) * (1.0 / 9007199254740992.0); }
const uint64_t STENCIL_SEED = 0xA5A5F00Dull;

template <class F>
void pfor_rows(i64 b, i64 e, F f) {
    i64 n = e - b;
    int nt = host_threads(n * 8);
    if (nt <= 1 || n < 1024) {
        f(b, e);
        return;
    }
    std::vector<std::thread> th;
    i64 chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        i64 lo = b + t * chunk, hi = std::min(e, lo + chunk);
        if (lo < hi) th.emplace_back([=]() { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

struct Grid {
    i64 nx, ny, nz;
};
inline int stencil_row_len(int kind, Grid G, i64 g) {
    if (kind == HPCLA_SYNTH_LAPLACE2D_5PT) {
        i64 ix = g % G.nx, iy = g / G.nx;
        return 1 + (ix > 0) + (ix < G.nx - 1) + (iy > 0) + (iy < G.ny - 1);
    }
    i64 ix = g % G.nx, iy = (g / G.nx) % G.ny, iz = g / (G.nx * G.ny);
    if (kind == HPCLA_SYNTH_POISSON3D_7PT) return 1 + (ix > 0) + (ix < G.nx - 1) + (iy > 0) + (iy < G.ny - 1) + (iz > 0) + (iz < G.nz - 1);
    int cx = 1 + (ix > 0) + (ix < G.nx - 1), cy = 1 + (iy > 0) + (iy < G.ny - 1), cz = 1 + (iz > 0) + (iz < G.nz - 1);
    return cx * cy * cz;
}

template <class T> struct Val;
template <> struct Val<float> { static float make(double re, double) { return (float)re; } };
template <> struct Val<double> { static double make(double re, double) { return re; } };
struct c128 { double re, im; };
template <> struct Val<c128> { static c128 make(double re, double im) { return c128{re, im}; } };

template <class T, class Ti>
void stencil_fill_rows(int kind, Grid G, i64 row_begin, i64 lo, i64 hi, const Ti* rowptr, Ti* cols, T* vals) {
    const i64 NX = G.nx, NY = G.ny, NZ = G.nz, NXY = G.nx * G.ny;
    for (i64 g = lo; g < hi; ++g) {
        i64 k = (i64)rowptr[g - row_begin] - 1;
        if (kind == HPCLA_SYNTH_LAPLACE2D_5PT) {
            i64 ix = g % NX, iy = g / NX;
            if (iy > 0) cols[k] = (Ti)(g - NX + 1), vals[k++] = Val<T>::make(-1, 0);
            if (ix > 0) cols[k] = (Ti)(g - 1 + 1), vals[k++] = Val<T>::make(-1, 0);
            cols[k] = (Ti)(g + 1), vals[k++] = Val<T>::make(4, 0);
            if (ix < NX - 1) cols[k] = (Ti)(g + 1 + 1), vals[k++] = Val<T>::make(-1, 0);
            if (iy < NY - 1) cols[k] = (Ti)(g + NX + 1), vals[k++] = Val<T>::make(-1, 0);
        } else if (kind == HPCLA_SYNTH_POISSON3D_7PT) {
            i64 ix = g % NX, iy = (g / NX) % NY, iz = g / NXY;
            if (iz > 0) cols[k] = (Ti)(g - NXY + 1), vals[k++] = Val<T>::make(-1, 0);
            if (iy > 0) cols[k] = (Ti)(g - NX + 1), vals[k++] = Val<T>::make(-1, 0);
            if (ix > 0) cols[k] = (Ti)(g), vals[k++] = Val<T>::make(-1, 0);
            cols[k] = (Ti)(g + 1), vals[k++] = Val<T>::make(6, 0);
            if (ix < NX - 1) cols[k] = (Ti)(g + 2), vals[k++] = Val<T>::make(-1, 0);
            if (iy < NY - 1) cols[k] = (Ti)(g + NX + 1), vals[k++] = Val<T>::make(-1, 0);
            if (iz < NZ - 1) cols[k] = (Ti)(g + NXY + 1), vals[k++] = Val<T>::make(-1, 0);
        } else {
            i64 ix = g % NX, iy = (g / NX) % NY, iz = g / NXY;
            int d = 0;
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx, ++d) {
                        i64 jx = ix + dx, jy = iy + dy, jz = iz + dz;
                        if (jx < 0 || jx >= NX || jy < 0 || jy >= NY || jz < 0 || jz >= NZ) continue;
                        uint64_t key = (uint64_t)(g * 27 + d);
                        double re = (d == 13 ? 26.0 : -1.0) + 0.1 * (2.0 * unit(STENCIL_SEED, 2 * key) - 1.0);
                        double im = 0.1 * (2.0 * unit(STENCIL_SEED, 2 * key + 1) - 1.0);
                        cols[k] = (Ti)(jx + NX * jy + NXY * jz + 1);
                        vals[k++] = Val<T>::make(re, im);
                    }
        }
    }
}

template <class T, class Ti>
int stencil_fill_typed(int kind, Grid N, i64 rb, i64 re, Ti* rowptr, Ti* cols, T* vals) {
    i64 acc = 1;
    for (i64 g = rb; g < re; ++g) {
        rowptr[g - rb] = (Ti)acc;
        acc += stencil_row_len(kind, N, g);
    }
    rowptr[re - rb] = (Ti)acc;
    pfor_rows(rb, re, [&](i64 lo, i64 hi) { stencil_fill_rows<T, Ti>(kind, N, rb, lo, hi, rowptr, cols, vals); });
    return HPCLA_OK;
}

inline i64 powerlaw_len(i64 n, uint64_t seed, i64 max_len, i64 g) {
    double u = unit(seed, (uint64_t)g);
    double L = std::floor(7.0 * std::pow(1.0 - u, -1.0 / 1.5));
    i64 len = L > 9.0e18 ? max_len : (i64)L;
    if (len > max_len) len = max_len;
    if (len > n) len = n;
    return len;
}

template <class T, class Ti>
int powerlaw_fill_typed(i64 n, uint64_t seed, i64 max_len, i64 rb, i64 re, Ti* rowptr, Ti* cols, T* vals) {
    i64 acc = 1;
    for (i64 g = rb; g < re; ++g) {
        rowptr[g - rb] = (Ti)acc;
        acc += powerlaw_len(n, seed, max_len, g);
    }
    rowptr[re - rb] = (Ti)acc;
    pfor_rows(rb, re, [&](i64 lo, i64 hi) {
        for (i64 g = lo; g < hi; ++g) {
            i64 k0 = (i64)rowptr[g - rb] - 1, L = (i64)rowptr[g - rb + 1] - 1 - k0;
            uint64_t rs = splitmix64(seed ^ (0xC0FFEEull + (uint64_t)g * 0x9E3779B97F4A7C15ull));
            for (i64 k = 0; k < L; ++k) {
                double u = unit(rs, 2 * (uint64_t)k);
                i64 c = (i64)std::floor(((double)k + u) * (double)n / (double)L);
                i64 lo_c = (i64)std::ceil((double)k * (double)n / (double)L);  // keep strictly inside the stratum
                if (c < lo_c) c = lo_c;
                if (c >= n) c = n - 1;
                cols[k0 + k] = (Ti)(c + 1);
                double v = 2.0 * unit(rs, 2 * (uint64_t)k + 1) - 1.0;
                vals[k0 + k] = Val<T>::make(v, 2.0 * unit(rs ^ 0x5555ull, (uint64_t)k) - 1.0);
            }
            for (i64 k = 1; k < L; ++k)  // float rounding at stratum edges: enforce strict ascent
                if (cols[k0 + k] <= cols[k0 + k - 1]) cols[k0 + k] = cols[k0 + k - 1] + 1;
        }
    });
    return HPCLA_OK;
}
}  // namespace

extern "C" const char* hpcla_synth_last_error(void) { return g_error.c_str(); }
extern "C" void hpcla_synth_set_threads(int n) { g_max_threads = n < 0 ? 0 : n; }

extern "C" int64_t hpcla_synth_stencil_rows(int kind, int64_t nx, int64_t ny, int64_t nz) { return kind == HPCLA_SYNTH_LAPLACE2D_5PT ? nx * ny : nx * ny * nz; }

extern "C" int64_t hpcla_synth_stencil_nnz(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t rb, int64_t re) {
    Grid N{nx, ny, nz};
    i64 acc = 0;
    for (i64 g = rb; g < re; ++g) acc += stencil_row_len(kind, N, g);
    return acc;
}

#define SYNTH_DISPATCH(FN, ...)                                                                                   \
    do {                                                                                                          \
        if (dtype == HPCLA_F32 && itype == HPCLA_I32) return FN<float, int32_t>(__VA_ARGS__, (int32_t*)rowptr, (int32_t*)global_cols, (float*)nzval);   \
        if (dtype == HPCLA_F32 && itype == HPCLA_I64) return FN<float, int64_t>(__VA_ARGS__, (int64_t*)rowptr, (int64_t*)global_cols, (float*)nzval);   \
        if (dtype == HPCLA_F64 && itype == HPCLA_I32) return FN<double, int32_t>(__VA_ARGS__, (int32_t*)rowptr, (int32_t*)global_cols, (double*)nzval); \
        if (dtype == HPCLA_F64 && itype == HPCLA_I64) return FN<double, int64_t>(__VA_ARGS__, (int64_t*)rowptr, (int64_t*)global_cols, (double*)nzval); \
        if (dtype == HPCLA_C128 && itype == HPCLA_I32) return FN<c128, int32_t>(__VA_ARGS__, (int32_t*)rowptr, (int32_t*)global_cols, (c128*)nzval);    \
        if (dtype == HPCLA_C128 && itype == HPCLA_I64) return FN<c128, int64_t>(__VA_ARGS__, (int64_t*)rowptr, (int64_t*)global_cols, (c128*)nzval);    \
        return fail(HPCLA_ERR_ARG, "synth: unknown dtype/itype");                                                 \
    } while (0)

extern "C" int hpcla_synth_stencil_fill(int kind, int64_t nx, int64_t ny, int64_t nz, int dtype, int itype, int64_t rb, int64_t re, void* rowptr, void* global_cols,
                                        void* nzval) {
    if (kind < 0 || kind > 2 || nx < 1 || ny < 1 || nz < 1 || rb < 0 || re < rb || re > hpcla_synth_stencil_rows(kind, nx, ny, nz))
        return fail(HPCLA_ERR_ARG, "hpcla_synth_stencil_fill: bad arguments");
    Grid N{nx, ny, nz};
    SYNTH_DISPATCH(stencil_fill_typed, kind, N, rb, re);
}

extern "C" int64_t hpcla_synth_powerlaw_nnz(int64_t n, uint64_t seed, int64_t max_len, int64_t rb, int64_t re) {
    i64 acc = 0;
    for (i64 g = rb; g < re; ++g) acc += powerlaw_len(n, seed, max_len, g);
    return acc;
}
extern "C" int hpcla_synth_powerlaw_fill(int64_t n, uint64_t seed, int64_t max_len, int dtype, int itype, int64_t rb, int64_t re, void* rowptr, void* global_cols,
                                         void* nzval) {
    if (n < 1 || rb < 0 || re < rb || re > n || max_len < 1) return fail(HPCLA_ERR_ARG, "hpcla_synth_powerlaw_fill: bad arguments");
    SYNTH_DISPATCH(powerlaw_fill_typed, n, seed, max_len, rb, re);
}

extern "C" int hpcla_synth_vector(int dtype, uint64_t seed, int64_t b, int64_t e, void* out) {
    if (b < 0 || e < b || !out) return fail(HPCLA_ERR_ARG, "hpcla_synth_vector: bad arguments");
    if (dtype == HPCLA_F32) { float* o = (float*)out; pfor_rows(b, e, [&](i64 lo, i64 hi) { for (i64 g = lo; g < hi; ++g) o[g - b] = (float)(2.0 * unit(seed, (uint64_t)g) - 1.0); }); }
    else if (dtype == HPCLA_F64) { double* o = (double*)out; pfor_rows(b, e, [&](i64 lo, i64 hi) { for (i64 g = lo; g < hi; ++g) o[g - b] = 2.0 * unit(seed, (uint64_t)g) - 1.0; }); }
    else if (dtype == HPCLA_C128) { c128* o = (c128*)out; pfor_rows(b, e, [&](i64 lo, i64 hi) { for (i64 g = lo; g < hi; ++g) o[g - b] = c128{2.0 * unit(seed, (uint64_t)g) - 1.0, 2.0 * unit(seed ^ 0xABCDEF12345ull, (uint64_t)g) - 1.0}; }); }
    else return fail(HPCLA_ERR_ARG, "hpcla_synth_vector: unknown dtype");
    return HPCLA_OK;
}
