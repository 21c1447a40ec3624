// gather_probe.cu — how fast can a B200 do fully scattered 4-byte gathers (the bound of the power-law multiply)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu && tools/gather_probe
// x: 20 M floats (80 MB, the size of x in BASELINE config 4); idx: 2^27 uniformly random indices streamed with 128-bit
// loads.  Variants: plain ld.global.nc, the same with an L2 evict-last hint, 16 gathers in flight per lane instead of 8,
// cp.async (LDGSTS) 4-byte gathers into shared memory, half of the lanes masked off per instruction, and a dependent
// "sorted within a warp" case.  Prints gathers per second and per SM clock.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e_));                     \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ float ldg_hint(const float* p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}

template <int MODE, int PER>  // PER: groups of 4 indices per lane per iteration
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ x, const int4* __restrict__ idx4, long long n4, float* __restrict__ out) {
    __shared__ float stage[256 * 4 * PER];
    uint64_t pol = 0;
    if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * 256 * PER;
    for (long long i = (long long)blockIdx.x * 256 * PER + threadIdx.x; i < n4; i += stride) {
        int4 c[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) c[u] = (i + (long long)u * 256 < n4) ? __ldcs(idx4 + i + (long long)u * 256) : make_int4(0, 0, 0, 0);
        if (MODE == 3) {  // cp.async 4-byte gathers into shared memory
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int cc[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&stage[(u * 4 + k) * 256 + threadIdx.x]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(x + cc[k]) : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 4 * PER; ++k) acc += stage[k * 256 + threadIdx.x];
        } else if (MODE == 4) {  // half the lanes per instruction (two instructions cover the warp)
            float v[4 * PER];
            const bool even = (threadIdx.x & 1) == 0;
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int cc[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a = 0.f, b = 0.f;
                    if (even) a = __ldg(x + cc[k]);
                    if (!even) b = __ldg(x + cc[k]);
                    v[u * 4 + k] = a + b;
                }
            }
#pragma unroll
            for (int k = 0; k < 4 * PER; ++k) acc += v[k];
        } else {
            float v[4 * PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                v[u * 4 + 0] = MODE == 1 ? ldg_hint(x + c[u].x, pol) : __ldg(x + c[u].x);
                v[u * 4 + 1] = MODE == 1 ? ldg_hint(x + c[u].y, pol) : __ldg(x + c[u].y);
                v[u * 4 + 2] = MODE == 1 ? ldg_hint(x + c[u].z, pol) : __ldg(x + c[u].z);
                v[u * 4 + 3] = MODE == 1 ? ldg_hint(x + c[u].w, pol) : __ldg(x + c[u].w);
            }
#pragma unroll
            for (int k = 0; k < 4 * PER; ++k) acc += v[k];
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

static uint64_t splitmix(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <int MODE, int PER>
static void run(const char* name, const float* x, const int* idx, long long n, float* out, int ctas_per_sm, int sms, double mhz) {
    const long long n4 = n / 4;
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) gather_kernel<MODE, PER><<<grid, 256>>>(x, (const int4*)idx, n4, out);
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; ++r) gather_kernel<MODE, PER><<<grid, 256>>>(x, (const int4*)idx, n4, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double gps = (double)n / (ms * 1e-3);
    printf("%-44s ctas/sm %d  %8.3f ms  %7.2f G gathers/s  %.3f gathers/clk/SM (at %.0f MHz)  idx stream %.0f GB/s\n", name, ctas_per_sm, ms, gps / 1e9,
           gps / sms / (mhz * 1e6), mhz, 4.0 * n / (ms * 1e-3) / 1e9);
}

int main(int argc, char** argv) {
    const long long nx = 20000000, n = argc > 1 ? atoll(argv[1]) : (1ll << 27);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, max %.0f MHz, L2 %d MB; x = %lld floats, %lld gathers per launch\n", prop.name, prop.multiProcessorCount, mhz, prop.l2CacheSize >> 20, nx, n);
    std::vector<int> h((size_t)n);
    uint64_t s = 42;
    for (long long i = 0; i < n; ++i) h[(size_t)i] = (int)(splitmix(s) % (uint64_t)nx);
    float *x, *out;
    int* idx;
    CK(cudaMalloc(&x, nx * 4));
    CK(cudaMemset(x, 0, nx * 4));
    CK(cudaMalloc(&out, 16));
    CK(cudaMalloc(&idx, n * 4));
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    const int sms = prop.multiProcessorCount;
    for (int c : {4, 8}) {
        run<0, 2>("ld.global.nc, 8 in flight per lane", x, idx, n, out, c, sms, mhz);
        run<0, 4>("ld.global.nc, 16 in flight per lane", x, idx, n, out, c, sms, mhz);
        run<1, 2>("ld.global.nc + L2 evict_last, 8 per lane", x, idx, n, out, c, sms, mhz);
        run<1, 4>("ld.global.nc + L2 evict_last, 16 per lane", x, idx, n, out, c, sms, mhz);
        run<3, 2>("cp.async 4 B -> shared, 8 per lane", x, idx, n, out, c, sms, mhz);
        run<3, 4>("cp.async 4 B -> shared, 16 per lane", x, idx, n, out, c, sms, mhz);
        run<4, 2>("ld.global.nc, half the lanes per instr", x, idx, n, out, c, sms, mhz);
    }
    // sorted within each group of 32 consecutive indices (what a warp sees if a long row's ascending columns are spread over lanes)
    for (long long i = 0; i + 128 <= n; i += 128) std::sort(h.begin() + i, h.begin() + i + 128);
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 2>("ld.global.nc, ascending within 128 (no locality)", x, idx, n, out, 8, sms, mhz);
    // small x (fits L1/L2 trivially): the pure issue / L1TEX rate
    for (long long i = 0; i < n; ++i) h[(size_t)i] &= 0xFFFF;
    CK(cudaMemcpy(idx, h.data(), n * 4, cudaMemcpyHostToDevice));
    run<0, 2>("ld.global.nc, x = 256 KB (L2 hits only)", x, idx, n, out, 8, sms, mhz);
    return 0;
}
