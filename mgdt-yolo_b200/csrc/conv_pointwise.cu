// Pointwise (1x1, stride 1) convolution for NARROW layers (Cin, Cout in {8, 16}, bf16) on the CUDA cores.
//
// These layers (the MSPA split branches) carry 16-32 bytes per pixel in and out and 64-256
// MACs per pixel: the arithmetic is free, the tensor-core kernel's per-tile pipeline (TMEM round trip, padded N = 16
// columns, 128-row tiles) is pure overhead and measured 20-27 us where the HBM time is 2-4 us.  Here one thread owns
// one pixel: one or two 16-byte loads, Cin x Cout FMAs against the fp32 weight matrix broadcast from shared memory,
// bias / activation / residual, one or two 16-byte stores -- fully coalesced, HBM-bound.
#include "common.cuh"

#include <algorithm>

namespace mgdt {

struct PwP {
    const __nv_bfloat16 *x, *w, *pre_add, *residual;
    const float* bias;
    __nv_bfloat16* y;
    unsigned M;
    int x_cs, y_cs, add_cs, res_cs, act;
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(256) conv_pointwise_kernel(PwP p) {
    pdl_trigger();
    __shared__ __align__(16) float sw[CIN][COUT];   // [ci][co]: a thread reads one row per input channel (broadcast)
    __shared__ float sb[COUT];
    // the weights are constant parameters: stage them before waiting on the producer of x
    for (int i = threadIdx.x; i < CIN * COUT; i += 256) {
        const int co = i / CIN, ci = i - co * CIN;   // OHWI: w[co][ci]
        sw[ci][co] = __bfloat162float(p.w[i]);
    }
    for (int i = threadIdx.x; i < COUT; i += 256) sb[i] = p.bias ? p.bias[i] : 0.f;
    __syncthreads();
    pdl_wait();
    for (unsigned pix = blockIdx.x * 256u + threadIdx.x; pix < p.M; pix += gridDim.x * 256u) {
        float xin[CIN];
#pragma unroll
        for (int c8 = 0; c8 < CIN; c8 += 8) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.x + (size_t)pix * p.x_cs + c8)), xin + c8);
            if (p.pre_add) {
                float a[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.pre_add + (size_t)pix * p.add_cs + c8)), a);
#pragma unroll
                for (int j = 0; j < 8; ++j) xin[c8 + j] += a[j];
                // the tensor-core path rounds the sum to bf16 before the MMA; do the same so both paths agree
#pragma unroll
                for (int j = 0; j < 8; ++j) xin[c8 + j] = __bfloat162float(__float2bfloat16_rn(xin[c8 + j]));
            }
        }
        float acc[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) acc[co] = sb[co];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
            for (int c4 = 0; c4 < COUT; c4 += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(&sw[ci][c4]);
                acc[c4] = fmaf(xin[ci], wv.x, acc[c4]);
                acc[c4 + 1] = fmaf(xin[ci], wv.y, acc[c4 + 1]);
                acc[c4 + 2] = fmaf(xin[ci], wv.z, acc[c4 + 2]);
                acc[c4 + 3] = fmaf(xin[ci], wv.w, acc[c4 + 3]);
            }
        }
        switch (p.act) {   // same SFU forms as the tcgen05 epilogue; the switch stays outside the unrolled loops
#define MGDT_ACT_CASE(A) case A: _Pragma("unroll") for (int co = 0; co < COUT; ++co) acc[co] = act_fast<A>(acc[co]); break;
            MGDT_ACT_CASE(MGDT_ACT_SILU)
            MGDT_ACT_CASE(MGDT_ACT_RELU)
            MGDT_ACT_CASE(MGDT_ACT_SIGMOID)
            MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
            MGDT_ACT_CASE(MGDT_ACT_GELU)
#undef MGDT_ACT_CASE
            default: break;
        }
#pragma unroll
        for (int c8 = 0; c8 < COUT; c8 += 8) {
            if (p.residual) {
                float r[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)pix * p.res_cs + c8)), r);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[c8 + j] += r[j];
            }
            uint4 o;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(acc[c8 + 2 * j], acc[c8 + 2 * j + 1]);
            *reinterpret_cast<uint4*>(p.y + (size_t)pix * p.y_cs + c8) = o;
        }
    }
}

static bool al16(const void* ptr, int cs) { return ptr == nullptr || ((((uintptr_t)ptr) & 15) == 0 && (cs & 7) == 0); }

bool conv2d_pointwise_supported(const mgdt_conv_args* a) {
    if (a->dtype != MGDT_BF16 || a->kh != 1 || a->kw != 1 || a->stride != 1 || a->pad != 0) return false;
    if (!((a->Cin == 8 || a->Cin == 16) && (a->Cout == 8 || a->Cout == 16))) return false;
    if (a->in_scale || a->pix_scale || a->in_relu) return false;
    if ((long long)a->N * a->H * a->W >= (1LL << 31)) return false;
    return al16(a->x, a->x_cs) && al16(a->y, a->y_cs) && al16(a->pre_add, a->add_cs) && al16(a->residual, a->res_cs);
}

int conv2d_pointwise(const mgdt_conv_args* a, cudaStream_t s) {
    PwP p;
    p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w; p.pre_add = (const __nv_bfloat16*)a->pre_add;
    p.residual = (const __nv_bfloat16*)a->residual; p.bias = a->bias; p.y = (__nv_bfloat16*)a->y;
    p.M = (unsigned)((long long)a->N * a->H * a->W);
    p.x_cs = a->x_cs; p.y_cs = a->y_cs; p.add_cs = a->add_cs; p.res_cs = a->res_cs; p.act = a->act;
    const int blocks = (int)std::min<long long>(((long long)p.M + 255) / 256, 148LL * 8);
#define MGDT_PW(CI, CO) launch_k(conv_pointwise_kernel<CI, CO>, dim3(blocks), dim3(256), 0, s, p)
    if (a->Cin == 8 && a->Cout == 8) MGDT_PW(8, 8);
    else if (a->Cin == 8 && a->Cout == 16) MGDT_PW(8, 16);
    else if (a->Cin == 16 && a->Cout == 8) MGDT_PW(16, 8);
    else MGDT_PW(16, 16);
#undef MGDT_PW
    MGDT_LAUNCH_CHECK("conv_pointwise");
    return 0;
}

}  // namespace mgdt
