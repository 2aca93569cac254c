// uint8 stem: predictor preprocessing (uint8 NCHW / 255, yolo/engine/predictor.py:115-130) fused into layer 0 of every
// config (Conv 3 -> Cout, k3 s2 p1 + BN + SiLU, nn/modules/conv.py:36-42) on warp-level tensor-core MMAs.
//
// Why not the tcgen05 kernel here: K = 27 and Cout = 16 make the layer pure data movement (39 MB of bytes in, 105 MB of
// bf16 out at B = 32: 22 us of HBM time) and the tcgen05 form spends ~530 thread instructions per output pixel gathering
// bytes into the UMMA operand layout (100 us, ncu: 2.0 IPC, issue-bound).  Here the bytes stay bytes until they are MMA
// fragments:
//   * a CTA stages the 2 TR + 1 input rows x 3 channels its TR output rows need as RAW uint8 rows in shared memory
//     (16-byte cp.async, rows are contiguous in NCHW; halo rows / margins zero-filled = the conv's padding);
//   * K is laid out as 12 groups (channel, ky) x 4 slots (input columns 2 ox - 2 .. 2 ox + 1; slot 0 and groups 9-11 carry
//     zero weights), so a thread's A fragment of mma.m16n8k16 -- two adjacent output pixels x two adjacent slots x two
//     groups per K step -- is two aligned 32-bit shared loads, one PRMT to pick its four bytes and one PRMT + one HSUB2
//     per register: byte b becomes the fp16 bit pattern 0x6400 | b = 1024 + b, minus 1024 = b exactly;
//   * weights are fp16 (scaled per output channel by a power of two so small BN-folded weights keep their mantissa;
//     1 / (255 * scale) is applied to the fp32 accumulator), packed once in B-fragment order and held in registers;
//   * output channels are permuted over the n-tiles so a thread's accumulators are 2 NT CONSECUTIVE channels of its
//     pixel: the NHWC store is one 8 / 16-byte store per pixel per thread, full 32-byte sectors per warp instruction.
#include "common.cuh"

#include <cuda_fp16.h>

#include <algorithm>

namespace mgdt {

constexpr int STEM_TR = 4;                      // output rows (= warps) per CTA
constexpr int STEM_RR = 2 * STEM_TR + 1;        // staged input rows per channel
constexpr int STEM_KSTEPS = 3;                  // 12 groups x 4 slots = 48 = 3 x k16

struct StemP {
    const uint8_t* src;        // (N, 3, H, W) uint8
    const uint2* wfrag;        // [3 k-steps][NT][32 lanes] {b0, b1} fp16 pairs
    const float* inv;          // [Cout] 1 / (255 * scale[co])
    const float* bias;         // [Cout] or null
    __nv_bfloat16* y;          // NHWC, channel stride y_cs
    int N, H, W, Ho, Wo, y_cs, act, P, tiles_y;
};

__device__ __forceinline__ void mma_f16_16816(float* d, const uint32_t* a, const uint2& b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// two bytes of v (selected by SEL) -> two exact fp16 integers
template <unsigned SEL> __device__ __forceinline__ uint32_t bytes_to_h2(uint32_t v) {
    const uint32_t m = __byte_perm(v, 0x64u, SEL);                // {0x6400 | lo, 0x6400 | hi} = {1024 + lo, 1024 + hi}
    const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), __half2half2(__ushort_as_half((unsigned short)0x6400)));
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t pack_bf2_(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

template <int NT, int ACT>   // NT = Cout / 8; ACT = compile-time activation or -1 (runtime p.act)
__global__ void __launch_bounds__(STEM_TR * 32) stem_mma_kernel(const StemP p) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n = blockIdx.x / p.tiles_y, oy0 = (blockIdx.x % p.tiles_y) * STEM_TR;
    const int iy0 = 2 * oy0 - 1, P = p.P, W = p.W;

    // ---- per-thread constants (weights: never written by a kernel, so before the dependency wait): B fragments, scale /
    // bias of the thread's 2 NT channels
    uint2 bf[STEM_KSTEPS][NT];
#pragma unroll
    for (int s = 0; s < STEM_KSTEPS; ++s)
#pragma unroll
        for (int j = 0; j < NT; ++j) bf[s][j] = __ldg(p.wfrag + (s * NT + j) * 32 + lane);
    float inv[2 * NT], bia[2 * NT];
#pragma unroll
    for (int j = 0; j < 2 * NT; ++j) {   // SiLU = h + h tanh(h), h = x / 2: the 1/2 is folded into scale and bias
        inv[j] = __ldg(p.inv + t * 2 * NT + j) * (ACT == MGDT_ACT_SILU ? 0.5f : 1.f);
        bia[j] = p.bias ? __ldg(p.bias + t * 2 * NT + j) * (ACT == MGDT_ACT_SILU ? 0.5f : 1.f) : 0.f;
    }
    // shared-memory row of the thread's six (channel, ky) groups: group = 2 i + (t >> 1), clamped (groups 9-11 have zero weights)
    const int oy = oy0 + warp;
    uint32_t roff[2 * STEM_KSTEPS];                                    // shared-space byte addresses
#pragma unroll
    for (int i = 0; i < 2 * STEM_KSTEPS; ++i) {
        const int grp = min(2 * i + (t >> 1), 8), c = grp / 3, ky = grp - 3 * c;
        roff[i] = (uint32_t)__cvta_generic_to_shared(smem) + (c * STEM_RR + 2 * warp + ky) * P + 12 + 4 * g;   // word holding columns 2 ox - 4 .. 2 ox - 1, ox = 2 g
    }
    const unsigned sel = (t & 1) ? 0x7654u : 0x5432u;                  // slots {2, 3} of pixel ox = slots {0, 1} of pixel ox + 1

    pdl_wait();   // the source may come from this library's LetterBox kernel; y may still be read by earlier consumers

    // ---- stage the raw rows
    {
        const int chunks = W >> 4, total = 3 * STEM_RR * chunks;
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
        for (int i = tid; i < total; i += STEM_TR * 32) {
            const int row = i / chunks, ch = i - row * chunks;
            const int c = row / STEM_RR, iy = iy0 + (row - c * STEM_RR);
            const uint32_t dst = sbase + row * P + 16 + ch * 16;
            if (iy >= 0 && iy < p.H) {
                const uint8_t* s = p.src + (((size_t)n * 3 + c) * p.H + iy) * W + ch * 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s) : "memory");
            } else {
                *reinterpret_cast<uint4*>(smem + row * P + 16 + ch * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        const int mwords = (P - W) >> 2;   // margins: 16 bytes left of column 0, everything right of column W - 1
        for (int i = tid; i < 3 * STEM_RR * mwords; i += STEM_TR * 32) {
            const int row = i / mwords, j = i - row * mwords;
            *reinterpret_cast<uint32_t*>(smem + row * P + (j < 4 ? j * 4 : W + j * 4)) = 0u;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (oy >= p.Ho) return;

    __nv_bfloat16* dst0 = p.y + (((size_t)n * p.Ho + oy) * p.Wo + 2 * g) * p.y_cs + t * 2 * NT;   // pixel ox = oxb + 2 g
    const size_t dstep = (size_t)16 * p.y_cs;
    const int act = p.act;
    for (int oxb = 0; oxb < p.Wo; oxb += 16, dst0 += dstep) {
        float acc[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int s = 0; s < STEM_KSTEPS; ++s) {
            uint32_t a[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t w0, w1;
                asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2 + 4];" : "=r"(w0), "=r"(w1) : "r"(roff[2 * s + h] + 2 * oxb));
                const uint32_t v = __byte_perm(w0, w1, sel);           // columns 2 ox - 2 + 2 (t & 1) .. + 3
                a[2 * h] = bytes_to_h2<0x4140u>(v);                    // pixel ox     (fragment row g)
                a[2 * h + 1] = bytes_to_h2<0x4342u>(v);                // pixel ox + 1 (fragment row g + 8)
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) mma_f16_16816(acc[j], a, bf[s][j]);
        }
        // ---- epilogue: thread holds channels t * 2 NT .. + 2 NT - 1 of pixels ox (acc[.][0..1]) and ox + 1 (acc[.][2..3])
        const int ox = oxb + 2 * g;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            uint32_t o[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                float v0 = fmaf(acc[j][2 * r], inv[2 * j], bia[2 * j]);
                float v1 = fmaf(acc[j][2 * r + 1], inv[2 * j + 1], bia[2 * j + 1]);
                if (ACT == MGDT_ACT_SILU) { v0 = fmaf(v0, tanh_fast(v0), v0); v1 = fmaf(v1, tanh_fast(v1), v1); }
                else if (ACT >= 0) { v0 = act_fast<ACT>(v0); v1 = act_fast<ACT>(v1); }
                else { v0 = act_fast_rt(v0, act); v1 = act_fast_rt(v1, act); }
                o[j] = pack_bf2_(v0, v1);
            }
            if (ox + r < p.Wo) {
                __nv_bfloat16* dst = dst0 + r * p.y_cs;
                if (NT % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < NT; j += 4) *reinterpret_cast<uint4*>(dst + 2 * j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < NT; j += 2) *reinterpret_cast<uint2*>(dst + 2 * j) = make_uint2(o[j], o[j + 1]);
                }
            }
        }
    }
}

// Pack the BN-folded fp32 weights (Cout, kp) with k = (ky * 3 + kx) * 3 + c into B-fragment order, fp16, scaled per output
// channel by a power of two; inv[co] = 1 / (255 * scale[co]).
__global__ void stem_mma_pack_kernel(const float* __restrict__ w, int kp, int Cout, uint2* __restrict__ frag, float* __restrict__ inv) {
    const int NT = Cout / 8;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= STEM_KSTEPS * NT * 32) return;
    const int lane = i & 31, j = (i >> 5) % NT, s = (i >> 5) / NT, g = lane >> 2, t = lane & 3;
    const int co = (g >> 1) * 2 * NT + 2 * j + (g & 1);   // fragment column g of n-tile j (see the kernel's epilogue)
    float mx = 0.f;
    for (int k = 0; k < 27; ++k) mx = fmaxf(mx, fabsf(w[(size_t)co * kp + k]));
    const float sc = mx > 0.f ? exp2f(floorf(log2f(8192.f / mx))) : 1.f;
    __half v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int kk = 2 * t + (e & 1) + 8 * (e >> 1), grp = 4 * s + (kk >> 2), slot = kk & 3;
        float x = 0.f;
        if (grp < 9 && slot >= 1) {
            const int c = grp / 3, ky = grp - 3 * c, kx = slot - 1;
            x = w[(size_t)co * kp + (ky * 3 + kx) * 3 + c] * sc;
        }
        v[e] = __float2half_rn(x);
    }
    uint2 o;
    o.x = (uint32_t)__half_as_ushort(v[0]) | ((uint32_t)__half_as_ushort(v[1]) << 16);
    o.y = (uint32_t)__half_as_ushort(v[2]) | ((uint32_t)__half_as_ushort(v[3]) << 16);
    frag[i] = o;
    if (s == 0 && t == 0) inv[co] = 1.f / (255.f * sc);
}

static int stem_pitch(int W, int Wo) {
    const int wo_pad = (Wo + 15) / 16 * 16;
    int P = 16 + std::max(W + 16, 2 * wo_pad);
    P = (P + 15) / 16 * 16;
    while (P % 128 < 32 || P % 128 > 96) P += 16;   // adjacent rows 8-24 banks apart: the two groups of a load do not collide
    return P;
}

template <int NT>
static int stem_launch(const StemP& p, size_t smem, cudaStream_t s) {
    const dim3 grid((unsigned)(p.N * p.tiles_y)), block(STEM_TR * 32);
    if (p.act == MGDT_ACT_SILU) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(stem_mma_kernel<NT, MGDT_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_k(stem_mma_kernel<NT, MGDT_ACT_SILU>, grid, block, smem, s, p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(stem_mma_kernel<NT, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_k(stem_mma_kernel<NT, -1>, grid, block, smem, s, p);
    }
    MGDT_LAUNCH_CHECK("stem_u8");
    return 0;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_stem_u8_supported(int C, int H, int W, int Cout, int y_cs) {
    if (C != 3 || H < 2 || W < 16 || (W & 15) || Cout < 16 || Cout > 80 || (Cout & 15) || (y_cs & 7) || y_cs < Cout) return 0;
    const int Wo = (W - 1) / 2 + 1;
    return (size_t)3 * STEM_RR * stem_pitch(W, Wo) <= 200 * 1024;
}

extern "C" size_t mgdt_stem_u8_packed_bytes(int Cout) {
    if (Cout < 16 || Cout > 80 || (Cout & 15)) return 0;
    return (size_t)STEM_KSTEPS * (Cout / 8) * 32 * sizeof(uint2) + (size_t)Cout * sizeof(float);
}

extern "C" int mgdt_stem_u8_pack(const float* w, int kp, int Cout, void* packed, void* stream) {
    MGDT_CHECK(w && packed && kp >= 27 && mgdt_stem_u8_packed_bytes(Cout) != 0, "stem_u8_pack: bad arguments");
    MGDT_CHECK(((uintptr_t)packed & 15) == 0, "stem_u8_pack: packed buffer must be 16-byte aligned");
    const int total = STEM_KSTEPS * (Cout / 8) * 32;
    launch_k(stem_mma_pack_kernel, dim3(cdiv(total, 128)), dim3(128), 0, (cudaStream_t)stream, w, kp, Cout, (uint2*)packed,
             (float*)((uint8_t*)packed + (size_t)total * sizeof(uint2)));
    MGDT_LAUNCH_CHECK("stem_u8_pack");
    return 0;
}

extern "C" int mgdt_stem_u8(const void* src, const void* packed, const float* bias, void* y, int y_cs, int N, int H, int W,
                            int Cout, int act, void* stream) {
    MGDT_CHECK(src && packed && y, "stem_u8: null pointer");
    MGDT_CHECK(N > 0 && mgdt_stem_u8_supported(3, H, W, Cout, y_cs), "stem_u8: unsupported shape %dx%d -> %d (cs %d)", H, W, Cout, y_cs);
    MGDT_CHECK(act >= MGDT_ACT_NONE && act <= MGDT_ACT_GELU, "stem_u8: bad act %d", act);
    MGDT_CHECK((((uintptr_t)src | (uintptr_t)y | (uintptr_t)packed) & 15) == 0, "stem_u8: source, output and packed weights must be 16-byte aligned");
    StemP p;
    p.src = (const uint8_t*)src; p.wfrag = (const uint2*)packed;
    p.inv = (const float*)((const uint8_t*)packed + (size_t)STEM_KSTEPS * (Cout / 8) * 32 * sizeof(uint2));
    p.bias = bias; p.y = (__nv_bfloat16*)y;
    p.N = N; p.H = H; p.W = W; p.Ho = (H - 1) / 2 + 1; p.Wo = (W - 1) / 2 + 1; p.y_cs = y_cs; p.act = act;
    p.P = stem_pitch(W, p.Wo); p.tiles_y = cdiv(p.Ho, STEM_TR);
    MGDT_CHECK((long long)N * p.tiles_y < (1LL << 31), "stem_u8: grid too large");
    const size_t smem = (size_t)3 * STEM_RR * p.P;
    cudaStream_t s = (cudaStream_t)stream;
    switch (Cout / 16) {   // NT = Cout / 8, even
        case 1: return stem_launch<2>(p, smem, s);
        case 2: return stem_launch<4>(p, smem, s);
        case 3: return stem_launch<6>(p, smem, s);
        case 4: return stem_launch<8>(p, smem, s);
        default: return stem_launch<10>(p, smem, s);
    }
}
