// TOODHead classification tail (nn/modules/head.py:519-521, 528) in one launch:
//
//     cls_prob = sigmoid(cls_prob_conv2(prob))          prob: (N, H, W, C1) after cls_prob_conv1 + ReLU, 3x3 pad 1 -> 1 channel
//     logits   = cv3(cls_feat * cls_prob)               cls_feat: (N, H, W, C2), 1x1 -> nc channels, written into raw[:, 4 reg_max:]
//
// As two tcgen05 convolutions these were a 16 -> 1 3x3 (18 us) and a 32 -> 2 1x1 with a per-pixel input scale (21 us) at
// B = 32: one and two useful output columns of a 128 x 16 UMMA tile.  The arithmetic is 200 MACs per pixel and the data
// 19 MB, so one thread per pixel on the CUDA cores is HBM-bound: nine 16-byte-chunk rows of `prob` (neighbouring threads
// share them through L1), the pixel's cls_feat row, fp32 weights broadcast from shared memory.  Rounding points are those
// of the two-kernel path: cls_prob and cls_feat * cls_prob are rounded to bf16, sums are fp32.
#include "common.cuh"

namespace mgdt {

struct TcP {
    const __nv_bfloat16 *prob, *w2, *feat, *w3;
    const float *b2, *b3;
    __nv_bfloat16* out;
    int N, H, W, C1, C2, prob_cs, feat_cs, out_cs;
};

__device__ __forceinline__ void tc_unpack8(const uint4& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

template <int NC>
__global__ void __launch_bounds__(128) tood_cls_kernel(const TcP p) {
    extern __shared__ __align__(16) float sw[];   // [9 * C1] conv2 weights (tap, channel), then [NC * C2] cv3 weights (C1, C2 % 8 == 0: 16-byte rows)
    pdl_trigger();
    float* s3 = sw + 9 * p.C1;
    for (int i = threadIdx.x; i < 9 * p.C1; i += 128) sw[i] = __bfloat162float(p.w2[i]);          // OHWI with O = 1
    for (int i = threadIdx.x; i < NC * p.C2; i += 128) s3[i] = __bfloat162float(p.w3[i]);
    __syncthreads();
    pdl_wait();
    const int q = blockIdx.x * 128 + threadIdx.x, n = blockIdx.y;
    if (q >= p.H * p.W) return;
    const int y = q / p.W, x = q - y * p.W;
    float a = p.b2 ? p.b2[0] : 0.f;
    // taps outside the image read a clamped (valid) address and are zeroed afterwards: no branch between the loads, so
    // the unrolled loop has all of a row's loads in flight before the first FMA
    const __nv_bfloat16* pn = p.prob + (size_t)n * p.H * p.W * p.prob_cs;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1, iyc = min(max(iy, 0), p.H - 1);
        const bool vy = iy == iyc;
        for (int c = 0; c < p.C1; c += 8) {
            uint4 v[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = x + kx - 1, ixc = min(max(ix, 0), p.W - 1);
                v[kx] = __ldg(reinterpret_cast<const uint4*>(pn + ((size_t)iyc * p.W + ixc) * p.prob_cs + c));
                if (!(vy && ix == ixc)) v[kx] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float f[8];
                tc_unpack8(v[kx], f);
                const float* w = sw + (ky * 3 + kx) * p.C1 + c;
                const float4 wa = *reinterpret_cast<const float4*>(w), wb = *reinterpret_cast<const float4*>(w + 4);
                a = fmaf(f[0], wa.x, a); a = fmaf(f[1], wa.y, a); a = fmaf(f[2], wa.z, a); a = fmaf(f[3], wa.w, a);
                a = fmaf(f[4], wb.x, a); a = fmaf(f[5], wb.y, a); a = fmaf(f[6], wb.z, a); a = fmaf(f[7], wb.w, a);
            }
        }
    }
    const float pr = __bfloat162float(__float2bfloat16_rn(act_fast<MGDT_ACT_SIGMOID>(a)));
    const size_t pix = ((size_t)n * p.H + y) * p.W + x;
    float acc[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) acc[k] = p.b3 ? p.b3[k] : 0.f;
    const __nv_bfloat16* fs = p.feat + pix * p.feat_cs;
    for (int c = 0; c < p.C2; c += 8) {
        float f[8];
        tc_unpack8(__ldg(reinterpret_cast<const uint4*>(fs + c)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16_rn(f[j] * pr));
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const float4 wa = *reinterpret_cast<const float4*>(s3 + k * p.C2 + c), wb = *reinterpret_cast<const float4*>(s3 + k * p.C2 + c + 4);
            float v = acc[k];
            v = fmaf(f[0], wa.x, v); v = fmaf(f[1], wa.y, v); v = fmaf(f[2], wa.z, v); v = fmaf(f[3], wa.w, v);
            v = fmaf(f[4], wb.x, v); v = fmaf(f[5], wb.y, v); v = fmaf(f[6], wb.z, v); v = fmaf(f[7], wb.w, v);
            acc[k] = v;
        }
    }
    __nv_bfloat16* o = p.out + pix * p.out_cs;
#pragma unroll
    for (int k = 0; k < NC; ++k) o[k] = __float2bfloat16_rn(acc[k]);
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_tood_cls_supported(int C1, int C2, int nc, int prob_cs, int feat_cs) {
    return C1 >= 8 && C1 <= 64 && (C1 & 7) == 0 && C2 >= 8 && C2 <= 256 && (C2 & 7) == 0 && nc >= 1 && nc <= 8 &&
           (prob_cs & 7) == 0 && (feat_cs & 7) == 0 && prob_cs >= C1 && feat_cs >= C2;
}

extern "C" int mgdt_tood_cls(const void* prob, int prob_cs, const void* w2, const float* b2, const void* feat, int feat_cs,
                             const void* w3, const float* b3, void* out, int out_cs, int N, int H, int W, int C1, int C2,
                             int nc, int dtype, void* stream) {
    MGDT_CHECK(prob && w2 && feat && w3 && out, "tood_cls: null pointer");
    MGDT_CHECK(dtype == MGDT_BF16, "tood_cls: bf16 only (the fp32 validation mode runs the two convolutions)");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && N <= 65535 && (long long)H * W < (1LL << 30) && out_cs >= nc, "tood_cls: bad shape");
    MGDT_CHECK(mgdt_tood_cls_supported(C1, C2, nc, prob_cs, feat_cs), "tood_cls: unsupported channels %d / %d / %d", C1, C2, nc);
    MGDT_CHECK((((uintptr_t)prob | (uintptr_t)feat) & 15) == 0, "tood_cls: inputs must be 16-byte aligned");
    TcP p;
    p.prob = (const __nv_bfloat16*)prob; p.w2 = (const __nv_bfloat16*)w2; p.feat = (const __nv_bfloat16*)feat;
    p.w3 = (const __nv_bfloat16*)w3; p.b2 = b2; p.b3 = b3; p.out = (__nv_bfloat16*)out;
    p.N = N; p.H = H; p.W = W; p.C1 = C1; p.C2 = C2; p.prob_cs = prob_cs; p.feat_cs = feat_cs; p.out_cs = out_cs;
    const dim3 grid(cdiv((long long)H * W, 128), N), block(128);
    const size_t smem = (size_t)(9 * C1 + nc * C2) * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    switch (nc) {
#define MGDT_TC(K) case K: launch_k(tood_cls_kernel<K>, grid, block, smem, s, p); break;
        MGDT_TC(1) MGDT_TC(2) MGDT_TC(3) MGDT_TC(4) MGDT_TC(5) MGDT_TC(6) MGDT_TC(7) MGDT_TC(8)
#undef MGDT_TC
    }
    MGDT_LAUNCH_CHECK("tood_cls");
    return 0;
}
