// Micro-benchmark: cp.async (LDGSTS) issue/throughput per warp on B200.  nvcc -arch=sm_100a -O3 -o cpasync cpasync.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
template <int VARIANT>
__global__ void k(const uint4* __restrict__ src, size_t stride_units, int n_per_thread, long long* out, int smem_units) {
    extern __shared__ uint4 sm[];
    const int tid = threadIdx.x;
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(sm);
    const uint4* g = src + (size_t)blockIdx.x * stride_units;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < n_per_thread; ++i) {
        const uint32_t idx = (uint32_t)(i * blockDim.x + tid);
        const uint32_t dst = s0 + (idx % smem_units) * 16u;
        if (VARIANT == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + idx) : "memory");
        if (VARIANT == 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(g + idx), "r"(16u) : "memory");
        if (VARIANT == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g + idx) : "memory");
        if (VARIANT == 3) { uint4 v = __ldg(g + idx); sm[idx % smem_units] = v; }
    }
    long long t1 = clock64();
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    long long t2 = clock64();
    if (tid == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
}
int main() {
    const int ctas = 148;
    const size_t per_cta_units = 1 << 16;  // 1 MiB per CTA
    uint4* src; long long* out;
    cudaMalloc(&src, ctas * per_cta_units * 16);
    cudaMemset(src, 1, ctas * per_cta_units * 16);
    cudaMalloc(&out, ctas * 2 * sizeof(long long));
    long long h[ctas * 2];
    const int smem_units = 8192;  // 128 KB
    auto run = [&](auto kern, const char* name, int threads, int n) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_units * 16);
        for (int rep = 0; rep < 2; ++rep) kern<<<ctas, threads, smem_units * 16>>>(src, per_cta_units, n, out, smem_units);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        double a = 0, b = 0;
        for (int i = 0; i < ctas; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
        a /= ctas; b /= ctas;
        const double bytes = (double)threads * n * 16;
        printf("%-28s threads=%4d n=%3d  issue %8.0f clk (%6.1f clk/instr/warp)  done %8.0f clk  -> %6.1f B/clk/SM  err=%s\n", name, threads, n, a,
               a / n, b, bytes / b, cudaGetErrorString(cudaGetLastError()));
    };
    for (int threads : {96, 256, 512}) {
        for (int n : {8, 43}) {
            run(k<0>, "cp.async.cg 16", threads, n);
            run(k<1>, "cp.async.cg 16 src-size", threads, n);
            run(k<2>, "cp.async.ca 16", threads, n);
            run(k<3>, "ldg+sts", threads, n);
        }
    }
    return 0;
}
