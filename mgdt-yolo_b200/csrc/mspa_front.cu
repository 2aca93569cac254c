// MSPA_C2f hierarchy front (nn/modules/block.py:248-262): the chain of 1x1 Conv+BN+SiLU branches
//
//     sp_0 = convs[0](spx[0]);  sp_i = convs[i](sp_{i-1} + spx[i])  (i = 1 .. g-2);   sp_in = sp_{g-2} + spx[g-1]
//
// is pointwise: every output pixel depends on the same pixel of x only.  Run as g-1 separate convolutions + one add it
// costs g launches that each read and write 16-128 bytes per pixel with 8-64 channels (42-75 us per MSPA block at
// B = 32 where the HBM time is 4-16 us).  Here ONE thread owns one pixel and walks the whole chain in registers: it
// reads the pixel's g*IW input channels once, writes the g-1 branch outputs into their slices of the concat buffer and
// the bottleneck input sp_in, and nothing else touches HBM.  Weights (fp32, [stage][ci][co]) are broadcast from shared
// memory; the inner product runs as packed fp32x2 FMAs.  Intermediate values are rounded to bf16 exactly where the
// unfused path stores / stages them (each branch output, each sum in front of the next conv), so both paths agree
// bit for bit.
#include "common.cuh"

#include <algorithm>

namespace mgdt {

struct MfP {
    const __nv_bfloat16* x;
    const float *w, *bias;
    __nv_bfloat16 *ycat, *ysp;
    unsigned M;
    int x_cs, y_cs, s_cs, act, nstage;
};

__device__ __forceinline__ unsigned long long mf_pk2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void mf_up2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long mf_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float mf_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <int IW> constexpr int mf_threads() { return IW >= 64 ? 128 : 256; }

template <int IW>
__global__ void __launch_bounds__(mf_threads<IW>()) mspa_front_kernel(MfP p) {
    pdl_trigger();
    constexpr int NT = mf_threads<IW>();
    constexpr int NV = IW / 8;   // 16-byte chunks per channel slice
    extern __shared__ __align__(16) float mf_smem[];
    float* sw = mf_smem;                              // [nstage][ci][co]
    float* sb = mf_smem + p.nstage * IW * IW;         // [nstage][co]
    // constant parameters: staged before waiting on the producer of x
    for (int i = threadIdx.x; i < p.nstage * IW * IW; i += NT) sw[i] = p.w[i];
    for (int i = threadIdx.x; i < p.nstage * IW; i += NT) sb[i] = p.bias[i];
    __syncthreads();
    pdl_wait();
    for (unsigned pix = blockIdx.x * (unsigned)NT + threadIdx.x; pix < p.M; pix += gridDim.x * (unsigned)NT) {
        const __nv_bfloat16* xr = p.x + (size_t)pix * p.x_cs;
        float cur[IW];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(xr) + v);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); cur[8 * v + 2 * j] = t.x; cur[8 * v + 2 * j + 1] = t.y; }
        }
#pragma unroll 1
        for (int st = 0; st < p.nstage; ++st) {
            uint4 nxt[NV];   // spx[st + 1]: in flight during the inner product
#pragma unroll
            for (int v = 0; v < NV; ++v) nxt[v] = __ldg(reinterpret_cast<const uint4*>(xr + (st + 1) * IW) + v);
            const float* ws = sw + st * IW * IW;
            unsigned long long acc[IW / 2];
#pragma unroll
            for (int c2 = 0; c2 < IW / 2; ++c2) acc[c2] = *reinterpret_cast<const unsigned long long*>(sb + st * IW + 2 * c2);
#pragma unroll
            for (int ci = 0; ci < IW; ++ci) {
                const unsigned long long xx = mf_pk2(cur[ci], cur[ci]);
#pragma unroll
                for (int c4 = 0; c4 < IW; c4 += 4) {
                    const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(ws + ci * IW + c4);
                    acc[c4 / 2] = mf_fma2(xx, wv.x, acc[c4 / 2]);
                    acc[c4 / 2 + 1] = mf_fma2(xx, wv.y, acc[c4 / 2 + 1]);
                }
            }
            float o[IW];
#pragma unroll
            for (int c2 = 0; c2 < IW / 2; ++c2) mf_up2(acc[c2], o[2 * c2], o[2 * c2 + 1]);
            switch (p.act) {   // same SFU forms as the conv epilogues; the switch stays outside the unrolled loops
#define MGDT_ACT_CASE(A) case A: _Pragma("unroll") for (int co = 0; co < IW; ++co) o[co] = act_fast<A>(o[co]); break;
                MGDT_ACT_CASE(MGDT_ACT_SILU)
                MGDT_ACT_CASE(MGDT_ACT_RELU)
                MGDT_ACT_CASE(MGDT_ACT_SIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_GELU)
#undef MGDT_ACT_CASE
                default: break;
            }
            __nv_bfloat16* yr = p.ycat + (size_t)pix * p.y_cs + st * IW;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                uint4 q;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(o[8 * v + 2 * j], o[8 * v + 2 * j + 1]);
                *(reinterpret_cast<uint4*>(yr) + v) = q;
                // next stage input: bf16(sp_st) + spx[st + 1], rounded to bf16 (what the unfused loaders stage)
                const __nv_bfloat162* a = reinterpret_cast<const __nv_bfloat162*>(&nxt[v]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 s = __bfloat1622float2(h[j]), t = __bfloat1622float2(a[j]);
                    cur[8 * v + 2 * j] = mf_round(s.x + t.x);
                    cur[8 * v + 2 * j + 1] = mf_round(s.y + t.y);
                }
            }
        }
        __nv_bfloat16* sr = p.ysp + (size_t)pix * p.s_cs;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            uint4 q;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(cur[8 * v + 2 * j], cur[8 * v + 2 * j + 1]);
            *(reinterpret_cast<uint4*>(sr) + v) = q;
        }
    }
}

template <int IW>
static int mspa_front_launch(const MfP& p, cudaStream_t s) {
    constexpr int NT = mf_threads<IW>();
    const size_t smem = (size_t)p.nstage * (IW * IW + IW) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(mspa_front_kernel<IW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-EIO, "mspa_front: smem attr: %s", cudaGetErrorString(e));
    const int blocks = (int)std::min<long long>(((long long)p.M + NT - 1) / NT, 148LL * 8);
    launch_k(mspa_front_kernel<IW>, dim3(blocks), dim3(NT), smem, s, p);
    MGDT_LAUNCH_CHECK("mspa_front");
    return 0;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_mspa_front_supported(int iw, int nstage) {
    return (iw == 8 || iw == 16 || iw == 32 || iw == 64) && nstage >= 1 && nstage <= 4 &&
           (size_t)nstage * (iw * iw + iw) * sizeof(float) <= 200 * 1024;
}

extern "C" int mgdt_mspa_front(const void* x, int x_cs, const float* w, const float* bias, int nstage, int iw, int act,
                               void* ycat, int y_cs, void* ysp, int s_cs, int N, int H, int W, int dtype, void* stream) {
    MGDT_CHECK(x && w && bias && ycat && ysp, "mspa_front: null pointer");
    MGDT_CHECK(dtype == MGDT_BF16, "mspa_front: bf16 only (the fp32 validation mode runs the branches as separate convs)");
    MGDT_CHECK(mgdt_mspa_front_supported(iw, nstage), "mspa_front: unsupported branch width %d / stage count %d", iw, nstage);
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && (long long)N * H * W < (1LL << 31), "mspa_front: bad shape");
    MGDT_CHECK(x_cs >= (nstage + 1) * iw && y_cs >= nstage * iw && s_cs >= iw, "mspa_front: channel stride too small");
    MGDT_CHECK(act >= MGDT_ACT_NONE && act <= MGDT_ACT_GELU, "mspa_front: bad act %d", act);
    MGDT_CHECK((((uintptr_t)x | (uintptr_t)ycat | (uintptr_t)ysp | (uintptr_t)w) & 15) == 0 && ((x_cs | y_cs | s_cs) & 7) == 0,
               "mspa_front: pointers must be 16-byte aligned and channel strides multiples of 8");
    MfP p;
    p.x = (const __nv_bfloat16*)x; p.w = w; p.bias = bias; p.ycat = (__nv_bfloat16*)ycat; p.ysp = (__nv_bfloat16*)ysp;
    p.M = (unsigned)((long long)N * H * W);
    p.x_cs = x_cs; p.y_cs = y_cs; p.s_cs = s_cs; p.act = act; p.nstage = nstage;
    cudaStream_t s = (cudaStream_t)stream;
    switch (iw) {
        case 8: return mspa_front_launch<8>(p, s);
        case 16: return mspa_front_launch<16>(p, s);
        case 32: return mspa_front_launch<32>(p, s);
        default: return mspa_front_launch<64>(p, s);
    }
}
