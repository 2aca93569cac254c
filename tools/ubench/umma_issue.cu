// Microbenchmark: how fast can ONE warp issue tcgen05.mma (M = 128, K = 16, bf16, N small) when the whole warp walks the
// loop in uniform control flow, all descriptor arithmetic is warp-uniform (kernel parameters + loop counters) and only
// the instruction itself sits under elect.sync (the CUTLASS pattern)?  Compared with the `if (elect) { loop }` form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_issue umma_issue.cu && ./umma_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc));
}

// mode 0: if (elect) { loop { mma } };  mode 1: loop { if (elect) mma } with uniform arithmetic
// Geometry of a 3x3 conv tile: taps shift the start address, MB row blocks (independent accumulators) innermost.
__constant__ int g_var;
struct Tab { unsigned long long a[160]; int n; };
template <int MODE>
__global__ void __launch_bounds__(128) rate(int N, int MB, int planes, int Wq, int lbo, int tiles, long long* out, const __grid_constant__ Tab tab) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t slot;
    const uint32_t base = (s_u32(sm) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 190 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (warp_u == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = base, b0 = base + 100 * 1024;
        const int halo = Wq + 1;
        const long long t0 = clock64();
        int n = 0;
        if (MODE == 0) {
            if (elect_one()) {
                for (int t = 0; t < tiles; ++t)
                    for (int tap = 0; tap < 9; ++tap)
                        for (int pp = 0; pp < planes / 2; ++pp) {
                            const uint32_t sh = (g_var & 2) ? 0u : (uint32_t)(halo + (tap / 3 - 1) * Wq + (tap % 3 - 1)) * 16u + (uint32_t)(2 * pp) * lbo;
                            const uint64_t bd = desc_plain(b0 + ((g_var & 1) ? 0u : (uint32_t)((tap * (planes / 2) + pp) * 2 * N * 16)), (uint32_t)N * 16u, 128u);
                            for (int mb = 0; mb < MB; ++mb)
                                mma(tmem + mb * N + ((g_var & 4) ? ((tap * (planes / 2) + pp) & 1) * MB * N : 0), desc_plain(a0 + sh + ((g_var & 8) ? 0 : mb * 2048), lbo, 128u), bd, idesc, (tap | pp) ? 1u : 0u);
                            n += MB;
                        }
            }
        } else if (MODE == 2) {
            // descriptor offsets from the kernel parameters (constant bank -> uniform registers), whole warp in the loop
            const bool lead = elect_one();
            const uint64_t abase = desc_plain(a0, 0, 128u), bbase = desc_plain(b0, (uint32_t)N * 16u, 128u);
            const uint32_t bstep = (uint32_t)(2 * N);
            for (int t = 0; t < tiles; ++t) {
                uint64_t bd = bbase;
                for (int i = 0; i < tab.n; ++i, bd += bstep) {
                    const uint64_t ad0 = abase + tab.a[i];
                    const uint32_t acc = i ? 1u : 0u;
                    for (int mb = 0; mb < MB; ++mb)
                        if (lead) mma(tmem + mb * N, ad0 + (uint64_t)(mb * 128), bd, idesc, acc);
                }
                n += tab.n * MB;
            }
        } else {
            const bool lead = elect_one();
            const uint64_t abase = desc_plain(a0, lbo, 128u), bbase = desc_plain(b0, (uint32_t)N * 16u, 128u);
            for (int t = 0; t < tiles; ++t) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int shp = halo + (tap / 3 - 1) * Wq + (tap % 3 - 1);      // uniform: parameters + unrolled constants
                    for (int pp = 0; pp < planes / 2; ++pp) {
                        const uint64_t ad0 = abase + (uint64_t)(shp + (2 * pp) * (lbo >> 4));
                        const uint64_t bd = bbase + (uint64_t)((tap * (planes / 2) + pp) * 2 * N);
                        for (int mb = 0; mb < MB; ++mb) {
                            if (lead) mma(tmem + mb * N, ad0 + (uint64_t)(mb * 128), bd, idesc, (tap | pp) ? 1u : 0u);
                        }
                        n += MB;
                    }
                }
            }
        }
        __syncwarp();
        if (elect_one()) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s_u32(&bar)), "r"(0u) : "memory");
            out[0] = clock64() - t0;
            out[1] = n;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* d;
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("cycles per 128xNx16 bf16 MMA, ONE issuing warp, 3x3 tap-shifted no-swizzle planes\n");
    for (int var : {0})
    for (int mode = 0; mode < 3; mode += 2)
        for (int planes : {4})
            for (int N : {32})
                for (int MB : {2, 4, 8}) {
                    if (MB * N * ((var & 4) ? 2 : 1) > 512) continue;
                    const int Wq = 82, lbo = ((MB * 128 + 2 * 83 + 15) / 8 * 8 + 8 + 1) * 16;
                    if (mode == 1) continue;
                    Tab tab; tab.n = 9 * planes / 2;
                    for (int tap = 0; tap < 9; ++tap) for (int pp = 0; pp < planes / 2; ++pp) tab.a[tap * (planes / 2) + pp] = (unsigned long long)((Wq + 1 + (tap / 3 - 1) * Wq + (tap % 3 - 1)) + 2 * pp * (lbo >> 4)) | ((unsigned long long)(lbo >> 4) << 16);
                    if ((size_t)planes * lbo > 100 * 1024) continue;
                    long long best = 1LL << 60, cnt = 0;
                    cudaMemcpyToSymbol(g_var, &var, 4);
                    for (int rep = 0; rep < 3; ++rep) {
                        if (mode == 0) rate<0><<<1, 128, 200 * 1024>>>(N, MB, planes, Wq, lbo, 8, d, tab);
                        else if (mode == 2) rate<2><<<1, 128, 200 * 1024>>>(N, MB, planes, Wq, lbo, 8, d, tab);
                        else rate<1><<<1, 128, 200 * 1024>>>(N, MB, planes, Wq, lbo, 8, d, tab);
                        long long h[2] = {0, 0};
                        if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                        if (h[0] < best) best = h[0];
                        cnt = h[1];
                    }
                    printf("var %d mode %d  Cin %2d N %3d MB %d | %7.1f cycles/MMA  (%lld MMAs)\n", var, mode, planes * 8, N, MB, (double)best / cnt, cnt);
                }
    return 0;
}
