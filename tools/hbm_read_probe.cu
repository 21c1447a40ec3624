// hbm_read_probe.cu — development probe: the read bandwidth a B200 sustains for (a) 128-bit LDG streaming reads and
// (b) cp.async.bulk (1-D TMA) global->shared reads shaped like the SpMV row-walk tiles.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_read_probe hbm_read_probe.cu && ./hbm_read_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void ldg_read(const int4* __restrict__ p, size_t n16, int* sink) {
    int acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        int4 v = __ldcs(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) *sink = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one CTA per tile of `tile_bytes`; one bulk copy (or `pieces` of them), wait, touch one word per thread
__global__ void tma_read(const char* __restrict__ p, size_t tile_bytes, int pieces, int* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    unsigned char* buf = smem + 16;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((uint32_t)tile_bytes) : "memory");
        const size_t piece = tile_bytes / pieces;
        for (int k = 0; k < pieces; ++k)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(buf + k * piece)),
                         "l"(p + (size_t)blockIdx.x * tile_bytes + k * piece), "r"((uint32_t)piece), "r"(s32(bar)), "l"(pol)
                         : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s32(bar)), "r"(0) : "memory");
    int v = reinterpret_cast<int*>(buf)[threadIdx.x];
    if (v == 0x12345678) *sink = v;
}

int main() {
    const size_t bytes = (size_t)1740 << 20;
    char* d;
    int* sink;
    cudaMalloc(&d, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time = [&](auto fn, const char* name) {
        for (int i = 0; i < 3; ++i) fn();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) fn();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-48s %8.1f us  %8.1f GB/s  (%s)\n", name, ms / 20 * 1e3, bytes / (ms / 20 * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        char name[64];
        snprintf(name, sizeof name, "ldg 128-bit streaming, %d CTAs x 256", blocks);
        time([&] { ldg_read<<<blocks, 256>>>((const int4*)d, bytes / 16, sink); }, name);
    }
    for (size_t tile : {(size_t)16384, (size_t)24576, (size_t)32768, (size_t)49152, (size_t)65536}) {
        for (int pieces : {1, 2}) {
            cudaFuncSetAttribute(tma_read, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile + 16);
            cudaFuncSetAttribute(tma_read, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            char name[64];
            snprintf(name, sizeof name, "tma bulk tiles of %zu KB x %d piece(s), 256 thr", tile >> 10, pieces);
            const unsigned grid = (unsigned)(bytes / tile);
            time([&] { tma_read<<<grid, 256, tile + 16>>>(d, tile, pieces, sink); }, name);
        }
    }
    return 0;
}
