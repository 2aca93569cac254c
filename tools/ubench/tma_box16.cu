// Microbenchmark: TMA tiled loads of NHWC bf16 activations as padded row blocks for the 3x3 implicit GEMM.
//   variant 0: one 4D box {8 ch, W+2, R+2, 1} per 8-channel plane (16-byte inner box: the no-swizzle K-major plane layout)
//   variant 1: one 4D box {C ch, W+2, R+2, 1} per tile (inner box = C*2 bytes)
// Reports GB/s of useful bytes (N*H*W*C*2) for persistent CTAs with a 3-stage ring and no consumer.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_box16 tma_box16.cu -lcuda && ./tma_box16
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32) k(const __grid_constant__ CUtensorMap map, int planes, int R, int rows_per_img_tiles, int ntiles,
                                         uint32_t stage_bytes, uint32_t plane_bytes, int variant, int S, uint32_t tx_bytes) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[8];
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % S;
        if (it >= S) {   // wait for the previous use of the stage to land
            const uint32_t par = ((it / S) - 1) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                             : "=r"(ok) : "r"(s_u32(&bars[s])), "r"(par) : "memory");
        }
        const int n = tile / rows_per_img_tiles, y0 = (tile % rows_per_img_tiles) * R - 1;
        const uint32_t bar = s_u32(&bars[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
        const uint32_t dst = s_u32(smem) + (uint32_t)s * stage_bytes;
        const int nld = variant == 0 ? planes : 1;
        for (int pl = 0; pl < nld; ++pl)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(dst + pl * plane_bytes), "l"(&map), "r"(pl * 8), "r"(-1), "r"(y0), "r"(n), "r"(bar) : "memory");
    }
    // drain
    for (int j = 0; j < S && j < it; ++j) {
        const int i2 = it - 1 - j, s = i2 % S;
        const uint32_t par = (i2 / S) & 1;
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(s_u32(&bars[s])), "r"(par) : "memory");
    }
}

int main() {
    const int N = 32;
    struct Cfg { int C, H, W, R; } cfgs[] = {{8, 160, 160, 3}, {16, 80, 80, 6}, {32, 80, 80, 6}, {64, 80, 80, 6}, {32, 40, 40, 12}, {64, 20, 20, 20}, {16, 160, 160, 3}};
    for (auto c : cfgs) {
        const size_t elems = (size_t)N * c.H * c.W * c.C;
        void* x;
        cudaMalloc(&x, elems * 2);
        cudaMemset(x, 1, elems * 2);
        for (int variant = 0; variant < 2; ++variant) {
            CUtensorMap map;
            cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)N};
            cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
            cuuint32_t box[4] = {(cuuint32_t)(variant == 0 ? 8 : c.C), (cuuint32_t)(c.W + 2), (cuuint32_t)(c.R + 2), 1}, es[4] = {1, 1, 1, 1};
            CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
            const int planes = c.C / 8;
            const uint32_t box_bytes = (uint32_t)(c.W + 2) * (c.R + 2) * 16;
            const uint32_t plane_bytes = (box_bytes + 127) / 128 * 128;
            const uint32_t stage_bytes = plane_bytes * planes, tx_bytes = box_bytes * planes;
            const int tpi = (c.H + c.R - 1) / c.R, ntiles = tpi * N;
            for (int ctas : {1, 2, 4}) {
                const int S = 3;
                const size_t smem = (size_t)S * stage_bytes + 1024;
                if (smem * ctas > 220 * 1024) continue;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                k<<<148 * ctas, 32, smem>>>(map, planes, c.R, tpi, ntiles, stage_bytes, plane_bytes, variant, S, tx_bytes);
                cudaEventRecord(e0);
                const int reps = 20;
                for (int i = 0; i < reps; ++i) k<<<148 * ctas, 32, smem>>>(map, planes, c.R, tpi, ntiles, stage_bytes, plane_bytes, variant, S, tx_bytes);
                cudaEventRecord(e1);
                cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                printf("C %3d %3dx%3d R %2d variant %d ctas/SM %d: %7.1f us  %7.1f GB/s useful  (%s)\n", c.C, c.H, c.W, c.R, variant, ctas,
                       ms * 1000 / reps, elems * 2.0 / (ms / reps * 1e-3) / 1e9, cudaGetErrorString(e));
            }
        }
        cudaFree(x);
    }
    return 0;
}
