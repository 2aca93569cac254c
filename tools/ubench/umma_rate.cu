// Microbenchmark: cycles per tcgen05.mma.cta_group::1.kind::f16 (M = 128, K = 16, bf16) issued back to back by one
// thread, as a function of N, the number of independent accumulators the instructions rotate over, and the A layout
// (no-swizzle K-major channel planes vs SWIZZLE_128B K-major rows; optional 16-byte misaligned start = a shifted tap).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(128) rate(int N, int chains, int layout, int shift, int b_sw, int iters, long long* out, int issuers, int lbo, int taps) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long bar[4];
    __shared__ uint32_t slot;
    const uint32_t base = (s_u32(sm) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar[w])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const int wq = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && wq < issuers) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = base + (shift ? 16u : 0u), b0 = base + 24 * 1024;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int k = i & 3;   // four K = 16 steps of a 64-channel block
            uint64_t ad, bd;
            if (layout == 0) {
                // taps: start shifted like the 9 taps of a 3x3 conv over a padded row pitch of 82 positions (16 B each)
                const int tp = taps ? (i % 9) : 0;
                const uint32_t sh = taps ? (uint32_t)(83 + (tp / 3 - 1) * 82 + (tp % 3 - 1)) * 16u : 0u;
                ad = desc_plain(a0 + sh + (uint32_t)(2 * (k & 1)) * (uint32_t)lbo, (uint32_t)lbo, 128u);
            }
            else ad = desc_sw128(a0 + (uint32_t)k * 32u);
            if (b_sw) bd = desc_sw128(b0 + (uint32_t)k * 32u);
            else bd = desc_plain(b0 + (uint32_t)(2 * k) * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
            const uint32_t d = tmem + (uint32_t)((wq * chains + (i % chains)) * N);
            const uint32_t acc = i >= chains ? 1u : 0u;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                         ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar[wq])) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s_u32(&bar[wq])), "r"(0u) : "memory");
        out[wq] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 252;
    printf("aggregate cycles per 128xNx16 bf16 MMA, 4 issuing threads, no-swizzle plane layout (A planes of P x 16 B)\n");
    struct V { const char* name; int shift, lbo, taps; };
    const V vs[] = {{"aligned start, plane stride 2304 B (18 x 128)", 0, 2304, 0}, {"start + 16 B", 1, 2304, 0},
                    {"plane stride 2320 B (odd x 16)", 0, 2320, 0}, {"3x3 tap shifts, stride 7056 B (441 x 16)", 0, 7056, 1},
                    {"3x3 tap shifts, stride 7040 B (55 x 128)", 0, 7040, 1}, {"3x3 tap shifts, stride 7168 B (56 x 128)", 0, 7168, 1}};
    for (const V& v : vs)
        for (int N : {16, 32, 64})
            for (int issuers : {1, 4}) {
                long long best = 1LL << 60;
                for (int rep = 0; rep < 3; ++rep) {
                    rate<<<1, 128, 64 * 1024>>>(N, 1, 0, v.shift, 0, iters, d, issuers, v.lbo, v.taps);
                    long long h[4] = {0, 0, 0, 0};
                    if (cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    long long mx = 0;
                    for (int w = 0; w < issuers; ++w) mx = h[w] > mx ? h[w] : mx;
                    if (mx < best) best = mx;
                }
                printf("%-48s N %3d issuers %d | aggregate %7.1f cycles/MMA\n", v.name, N, issuers, (double)best / iters / issuers);
            }
    return 0;
}
