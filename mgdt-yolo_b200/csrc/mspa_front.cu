// MSPA_C2f hierarchy front (nn/modules/block.py:248-262): the chain of 1x1 Conv+BN+SiLU branches
//
//     sp_0 = convs[0](spx[0]);  sp_i = convs[i](sp_{i-1} + spx[i])  (i = 1 .. g-2);   sp_in = sp_{g-2} + spx[g-1]
//
// is pointwise: every output pixel depends on the same pixel of x only.  Run as g-1 separate convolutions + one add it
// costs g launches that each read and write 16-128 bytes per pixel with 8-64 channels (42-75 us per MSPA block at
// B = 32 where the HBM time is 4-16 us).  Here a WARP owns 32 pixels and walks the whole chain in registers with
// warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 x bf16 -> fp32): the accumulator fragment of stage i, after
// bias / activation / bf16 rounding / + spx[i+1], IS the A fragment of stage i+1 (the two 8-column accumulator tiles
// 2k, 2k+1 of a row pair hold exactly the elements of A's k-block k), so nothing is exchanged between lanes and
// nothing but x, the branch outputs (into their slices of the concat buffer) and sp_in touches HBM.  The first
// version of this kernel (one thread per pixel, fp32 FMAs against weights broadcast from shared memory) was
// issue-bound at ~10 TFLOP/s for every width (26-30 us per launch, the MAC count is the same at all four levels);
// the tensor-core form needs ~1/3 (IW = 8) to ~1/7 (IW = 64) of its instructions.  The tile is far too small for
// tcgen05 (M = 128 rows x N = IW <= 64 with a TMEM round trip per stage): that path is what conv_umma2 is for.
// Weights are staged once per CTA in shared memory in B-fragment order (mgdt_mspa_front_pack).
#include "common.cuh"

#include <algorithm>

namespace mgdt {

struct MfP {
    const __nv_bfloat16* x;
    const uint2* wfrag;      // [nstage][IW/8 n-tiles][KB k-blocks][32 lanes] {b0, b1}
    const float* bias;       // [nstage][IW]
    __nv_bfloat16 *ycat, *ysp;
    unsigned M;
    int x_cs, y_cs, s_cs, act, nstage;
};

template <int IW> constexpr int mf_threads() { return IW >= 64 ? 64 : IW >= 32 ? 128 : 256; }

__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, const uint2& b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t add_bf2(uint32_t a, uint32_t b) {   // bf16(a + b) per half, one rounding
    const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

template <int IW>
__global__ void __launch_bounds__(mf_threads<IW>()) mspa_front_kernel(MfP p) {
    pdl_trigger();
    constexpr int NT = mf_threads<IW>();
    constexpr int KB = (IW + 15) / 16;   // 16-channel k-blocks (IW = 8: one, upper half zero)
    constexpr int NTL = IW / 8;          // 8-channel accumulator tiles
    extern __shared__ __align__(16) unsigned char mf_smem[];
    uint2* sfrag = reinterpret_cast<uint2*>(mf_smem);
    float* sb = reinterpret_cast<float*>(sfrag + p.nstage * NTL * KB * 32);
    // constant parameters: staged before waiting on the producer of x
    for (int i = threadIdx.x; i < p.nstage * NTL * KB * 32; i += NT) sfrag[i] = p.wfrag[i];
    for (int i = threadIdx.x; i < p.nstage * IW; i += NT) sb[i] = p.bias[i];
    __syncthreads();
    pdl_wait();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const unsigned wstep = gridDim.x * (unsigned)(NT / 32) * 32u;
    for (unsigned base = (blockIdx.x * (unsigned)(NT / 32) + (threadIdx.x >> 5)) * 32u; base < p.M; base += wstep) {
        // this lane's four pixel rows: (m-tile mt, half h) -> base + mt*16 + g + 8h
        const __nv_bfloat16* xr[2][2];
        bool ok[2][2];
        unsigned row[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                row[mt][h] = base + mt * 16 + g + 8 * h;
                ok[mt][h] = row[mt][h] < p.M;
                xr[mt][h] = p.x + (size_t)(ok[mt][h] ? row[mt][h] : 0u) * p.x_cs + 2 * t;
            }
        // slice `sl` of x in A-fragment layout: [mt][kb][j*2 + h] = channels kb*16 + 8j + 2t, +1 of row (mt, h)
        auto load_slice = [&](int sl, uint32_t (&X)[2][KB][4]) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint32_t v = 0u;
                            if (kb * 16 + 8 * j < IW && ok[mt][h])
                                v = __ldg(reinterpret_cast<const uint32_t*>(xr[mt][h] + sl * IW + kb * 16 + 8 * j));
                            X[mt][kb][j * 2 + h] = v;
                        }
        };
        uint32_t A[2][KB][4];
        load_slice(0, A);
#pragma unroll 1
        for (int st = 0; st < p.nstage; ++st) {
            uint32_t X[2][KB][4];   // spx[st + 1], in flight during the MMAs
            load_slice(st + 1, X);
            float D[2][NTL][4];
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
                const float2 b = *reinterpret_cast<const float2*>(sb + st * IW + nt * 8 + 2 * t);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) { D[mt][nt][0] = D[mt][nt][2] = b.x; D[mt][nt][1] = D[mt][nt][3] = b.y; }
            }
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt) {
                    const uint2 b = sfrag[((st * NTL + nt) * KB + kb) * 32 + lane];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) mma_bf16_16816(D[mt][nt], A[mt][kb], b);
                }
            switch (p.act) {   // same SFU forms as the conv epilogues; the switch stays outside the unrolled loops
#define MGDT_ACT_CASE(ACT)                                                                        \
    case ACT:                                                                                     \
        _Pragma("unroll") for (int mt = 0; mt < 2; ++mt)                                          \
        _Pragma("unroll") for (int nt = 0; nt < NTL; ++nt)                                        \
        _Pragma("unroll") for (int e = 0; e < 4; ++e) D[mt][nt][e] = act_fast<ACT>(D[mt][nt][e]); \
        break;
                MGDT_ACT_CASE(MGDT_ACT_SILU)
                MGDT_ACT_CASE(MGDT_ACT_RELU)
                MGDT_ACT_CASE(MGDT_ACT_SIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_GELU)
#undef MGDT_ACT_CASE
                default: break;
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        // sp_st rounded to bf16: stored into its concat slice, and + spx[st + 1] (rounded again, as the
                        // unfused loaders stage it) it becomes the A fragment of the next stage
                        const uint32_t v = pack_bf2(D[mt][nt][2 * h], D[mt][nt][2 * h + 1]);
                        if (ok[mt][h])
                            *reinterpret_cast<uint32_t*>(p.ycat + (size_t)row[mt][h] * p.y_cs + st * IW + nt * 8 + 2 * t) = v;
                        A[mt][nt >> 1][(nt & 1) * 2 + h] = add_bf2(v, X[mt][nt >> 1][(nt & 1) * 2 + h]);
                    }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (kb * 16 + 8 * j < IW && ok[mt][h])
                            *reinterpret_cast<uint32_t*>(p.ysp + (size_t)row[mt][h] * p.s_cs + kb * 16 + 8 * j + 2 * t) = A[mt][kb][j * 2 + h];
    }
}

// fp32 [nstage][ci][co] -> bf16 B fragments [nstage][nt][kb][lane]{b0, b1}: b_r = {W[kb*16 + 2t + 8r][nt*8 + g], W[.. + 1][..]}
__global__ void mspa_front_pack_kernel(const float* __restrict__ w, uint2* __restrict__ out, int nstage, int iw) {
    // no pdl_trigger(): the consumer stages the packed weights before its griddepcontrol.wait (see umma2_pack_kernel)
    pdl_wait();
    const int KB = (iw + 15) / 16, NTL = iw / 8;
    const int total = nstage * NTL * KB * 32;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int lane = i & 31, kb = (i >> 5) % KB, nt = (i / (32 * KB)) % NTL, st = i / (32 * KB * NTL);
        const int g = lane >> 2, t = lane & 3, co = nt * 8 + g;
        uint32_t r[2];
        for (int rr = 0; rr < 2; ++rr) {
            const int ci = kb * 16 + 2 * t + 8 * rr;
            const float lo = ci < iw ? w[((size_t)st * iw + ci) * iw + co] : 0.f;
            const float hi = ci + 1 < iw ? w[((size_t)st * iw + ci + 1) * iw + co] : 0.f;
            r[rr] = pack_bf2(lo, hi);
        }
        out[i] = make_uint2(r[0], r[1]);
    }
}

template <int IW>
static int mspa_front_launch(const MfP& p, cudaStream_t s) {
    constexpr int NT = mf_threads<IW>();
    constexpr int KB = (IW + 15) / 16, NTL = IW / 8;
    const size_t smem = (size_t)p.nstage * (NTL * KB * 32 * sizeof(uint2) + IW * sizeof(float));
    cudaError_t e = cudaFuncSetAttribute(mspa_front_kernel<IW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(-EIO, "mspa_front: smem attr: %s", cudaGetErrorString(e));
    int occ = 1;   // resident CTAs per SM: a grid of whole waves avoids a ragged second wave
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mspa_front_kernel<IW>, NT, smem);
    const long long need = ((long long)p.M + NT - 1) / NT;   // one 32-pixel group per warp
    const int blocks = (int)std::min<long long>(need, 148LL * std::max(occ, 1));
    launch_k(mspa_front_kernel<IW>, dim3(blocks), dim3(NT), smem, s, p);
    MGDT_LAUNCH_CHECK("mspa_front");
    return 0;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_mspa_front_supported(int iw, int nstage) {
    return (iw == 8 || iw == 16 || iw == 32 || iw == 64) && nstage >= 1 && nstage <= 8;
}

extern "C" size_t mgdt_mspa_front_packed_bytes(int iw, int nstage) {
    if (!mgdt_mspa_front_supported(iw, nstage)) return 0;
    return (size_t)nstage * (iw / 8) * ((iw + 15) / 16) * 32 * sizeof(uint2);
}

extern "C" int mgdt_mspa_front_pack(const float* w, int nstage, int iw, void* packed, void* stream) {
    MGDT_CHECK(w && packed && mgdt_mspa_front_supported(iw, nstage), "mspa_front_pack: bad arguments");
    MGDT_CHECK(((uintptr_t)packed & 15) == 0, "mspa_front_pack: packed buffer must be 16-byte aligned");
    const int total = nstage * (iw / 8) * ((iw + 15) / 16) * 32;
    launch_k(mspa_front_pack_kernel, dim3(cdiv(total, 256)), dim3(256), 0, (cudaStream_t)stream, w, (uint2*)packed, nstage, iw);
    MGDT_LAUNCH_CHECK("mspa_front_pack");
    return 0;
}

extern "C" int mgdt_mspa_front(const void* x, int x_cs, const void* w_packed, const float* bias, int nstage, int iw, int act,
                               void* ycat, int y_cs, void* ysp, int s_cs, int N, int H, int W, int dtype, void* stream) {
    MGDT_CHECK(x && w_packed && bias && ycat && ysp, "mspa_front: null pointer");
    MGDT_CHECK(dtype == MGDT_BF16, "mspa_front: bf16 only (the fp32 validation mode runs the branches as separate convs)");
    MGDT_CHECK(mgdt_mspa_front_supported(iw, nstage), "mspa_front: unsupported branch width %d / stage count %d", iw, nstage);
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && (long long)N * H * W < (1LL << 31) - 64, "mspa_front: bad shape");
    MGDT_CHECK(x_cs >= (nstage + 1) * iw && y_cs >= nstage * iw && s_cs >= iw, "mspa_front: channel stride too small");
    MGDT_CHECK(act >= MGDT_ACT_NONE && act <= MGDT_ACT_GELU, "mspa_front: bad act %d", act);
    MGDT_CHECK((((uintptr_t)x | (uintptr_t)ycat | (uintptr_t)ysp) & 3) == 0 && ((x_cs | y_cs | s_cs) & 1) == 0 &&
                   ((uintptr_t)w_packed & 15) == 0,
               "mspa_front: activations must be 4-byte aligned with even channel strides, packed weights 16-byte aligned");
    MfP p;
    p.x = (const __nv_bfloat16*)x; p.wfrag = (const uint2*)w_packed; p.bias = bias;
    p.ycat = (__nv_bfloat16*)ycat; p.ysp = (__nv_bfloat16*)ysp;
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
