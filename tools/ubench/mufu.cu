// MUFU throughput per SM sub-partition: tanh.approx vs ex2.approx vs rcp.approx vs FFMA2 polynomial (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters) {
    float v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[j]));
            if (OP == 4) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
        }
    }
    float s = 0;
    for (int j = 0; j < 8; ++j) s += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> void run(const char* name, int warps_per_sm) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<148, warps_per_sm * 32>>>(out, 16);
    cudaEventRecord(a);
    k<OP><<<148, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)iters * 8 * warps_per_sm;   // warp-instructions per SM
    printf("%-6s warps/SM %2d: %.2f ns per warp-instr per SM  (%.2f cycles @1.965GHz per SMSP-instr)\n", name, warps_per_sm,
           ms * 1e6 / ops, ms * 1e6 / ops * 1.965 * 4);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16}) {
        run<0>("tanh", w); run<1>("ex2", w); run<2>("rcp", w); run<4>("rsqrt", w); run<3>("ffma", w);
    }
    return 0;
}
