// Microbenchmark: warp-level mma.sync.m16n8k16 bf16 (legacy HMMA path) throughput per SM on sm_100a, as a function of
// resident warps and independent accumulator chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(int iters, float* out, long long* cyc) {
    float acc[CHAINS][4];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 12345.f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* o; long long* c;
    cudaMalloc(&o, 4); cudaMalloc(&c, 8 * 1024);
    const int iters = 2000;
    printf("mma.sync.m16n8k16 bf16 (2048 MAC): MAC / clk / SM, one CTA per SM\n");
    for (int warps : {4, 8, 16, 32})
        for (int chains : {1, 2, 4, 8}) {
            long long h = 0;
            if (chains == 1) k<1><<<148, warps * 32>>>(iters, o, c);
            if (chains == 2) k<2><<<148, warps * 32>>>(iters, o, c);
            if (chains == 4) k<4><<<148, warps * 32>>>(iters, o, c);
            if (chains == 8) k<8><<<148, warps * 32>>>(iters, o, c);
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            const double mac = (double)iters * chains * warps * 2048.0;
            printf("warps %2d chains %d | %8.1f MAC/clk/SM  (%.1f cycles per HMMA per warp)\n", warps, chains, mac / h, (double)h / (iters * chains));
        }
    return 0;
}
