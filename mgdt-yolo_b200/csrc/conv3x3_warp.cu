// 3x3 stride-1 convolutions with few channels (Cin = Cout in {8, 16, 32}: the Bottleneck pairs of MSPA_C2f / C2f,
// nn/modules/block.py:514-526, at 160^2 ... 40^2) on warp-level tensor-core MMAs.
//
// Why not the tcgen05 kernels here: these layers move 1.6-13 MB and need 0.2-0.9 GFLOP per batch of 32 -- 1-3 us of
// either roofline -- but took 14-28 us each on the persistent tcgen05 kernel: a 128-row UMMA tile with N = 8-32 columns
// leaves the tensor pipe idle behind its own set-up (TMEM allocation, barrier ring, descriptor table: ~3 us to the first
// MMA, ~2 us tail) and a per-unit epilogue cost that does not shrink with N.  A warp-level m16n8k16 chain has no set-up:
//   * a persistent CTA (4 warps, one output row each) walks strips of TH output rows, double-buffered: while it computes a
//     strip the next one's TH + 2 input rows are staged as they are (NHWC, zero-filling 16-byte
//     cp.async = the conv's padding), pixel pitch padded by 16 bytes so that the eight 16-byte rows of an ldmatrix tile
//     fall into disjoint banks;
//   * K is ordered (tap, channel); one ldmatrix.x4 per K = 16 step delivers the A fragment of 16 consecutive pixels of the
//     row (Cin = 8: a step is two taps, the lanes of matrices 2 / 3 address the second one; tap 9 has zero weights);
//   * B fragments: straight from the OHWI bf16 weights into registers (<= 36), or for 32 -> 32 the weight rows staged by
//     cp.async in fragment-column order and fetched with ldmatrix.x4, shared by two pixel tiles per step;
//   * output channels are permuted over the n-tiles so a thread owns 2 NT consecutive channels of its pixels: bias,
//     activation, residual (requested before the MMA chain) and the NHWC store are 4 / 8 / 16-byte accesses; all
//     per-tile addresses advance by pointer increments (no index arithmetic in the loop).
// Tried and removed: the same scheme for the stride-2 layers 16 -> 32 @320^2 / 32 -> 64 @160^2 (input rows split by column
// parity while staging): correct, but 72 / 54 us against 55 / 37 us on the TMA-fed tcgen05 kernel -- with 2 CTAs of 4 warps
// per SM (85 KB of staged rows each) a load-then-compute CTA cannot hide its own staging latency.
#include "common.cuh"

#include <algorithm>

namespace mgdt {

int g_conv3x3_warp = 2;   // mgdt_set_option("conv3x3_warp", v): 0 = these layers stay on the tcgen05 kernels, 1 = 8 -> 8 / 16 -> 16 only, 2 = also 32 -> 32
int g_conv3x3_warp_spc = 1;   // "conv3x3_warp_spc": target strips per persistent CTA (1 = one CTA per strip, no pipelining)

#ifndef MGDT_CW_TH
#define MGDT_CW_TH 4
#endif
constexpr int CW_TH = MGDT_CW_TH;   // output rows per CTA = warps per CTA (variant build -DMGDT_CW_TH=8: -1 % images/s in an in-box A/B)

struct CwP {
    const __nv_bfloat16 *x, *w, *res;
    const float* bias;
    __nv_bfloat16* y;
    int N, H, W, x_cs, y_cs, res_cs, act, TX, TWp, tiles_y, strips, nbuf;
};

__device__ __forceinline__ void mma_bf16_16816_(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf2w(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
template <int NW> struct CwVec;
template <> struct CwVec<1> { using T = uint32_t; };
template <> struct CwVec<2> { using T = uint2; };
template <> struct CwVec<4> { using T = uint4; };

template <int CIN> constexpr int cw_pitch() { return CIN == 8 ? 16 : CIN * 2 + 16; }   // bytes per staged pixel
template <int CIN> constexpr int cw_ksteps() { return CIN == 8 ? 5 : 9 * (CIN / 16); }
template <int CIN> constexpr int cw_wpitch() { return 9 * CIN * 2 + 16; }                // bytes per staged weight row (32 -> 32)
constexpr int CW_SLACK = 16;                                                             // staged pixels past the last row (second tile of a pair)

template <int CIN, int COUT, int ACT>
__global__ void __launch_bounds__(CW_TH * 32) conv3x3_warp_kernel(const CwP p) {
    constexpr int NT = COUT / 8, KS = cw_ksteps<CIN>(), PITCH = cw_pitch<CIN>(), CPP = CIN / 8, C16 = CIN >= 16 ? CIN / 16 : 1;
    constexpr bool B_REGS = KS * NT <= 18;
    constexpr int MT = B_REGS ? 1 : 2, WP = cw_wpitch<CIN>();
    using Vec = typename CwVec<NT>::T;
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int TWp = p.TWp;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bufsz = ((CW_TH + 2) * TWp + CW_SLACK) * PITCH;           // one staged strip; two of them (double buffer) when a CTA walks several strips
    const uint32_t wbase = sbase + p.nbuf * bufsz;                           // staged weight rows (32 -> 32 only)

    // ---- weights (never written by a kernel of the stream: before the dependency wait).  Column q of n-tile j is output
    // channel (q >> 1) * 2 NT + 2 j + (q & 1), so thread t's accumulator columns 2t, 2t+1 over j are channels t * 2 NT ...
    uint32_t bf[B_REGS ? KS : 1][B_REGS ? NT : 1][2];
    if (B_REGS) {
        const __nv_bfloat16* wrow[NT];
#pragma unroll
        for (int j = 0; j < NT; ++j) wrow[j] = p.w + (size_t)((g >> 1) * 2 * NT + 2 * j + (g & 1)) * 9 * CIN + 2 * t;
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {     // k = 16 s + 8 h + 2t, 2t + 1 of the (tap, channel) order; Cin = 8: tap 2 s + h
                    const int k = 16 * s + 8 * h;
                    bf[s][j][h] = (CIN == 8 && k >= 72) ? 0u : __ldg(reinterpret_cast<const uint32_t*>(wrow[j] + k));
                }
    } else {
        for (int i = tid; i < COUT * (9 * CIN / 8); i += CW_TH * 32) {           // row rho = 8 j + q holds channel (q >> 1) * 2 NT + 2 j + (q & 1)
            const int rho = i / (9 * CIN / 8), ch = i - rho * (9 * CIN / 8), j = rho >> 3, q = rho & 7;
            const int co = (q >> 1) * 2 * NT + 2 * j + (q & 1);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wbase + rho * WP + ch * 16), "l"(p.w + (size_t)co * 9 * CIN + ch * 8) : "memory");
        }
    }
    float bia[2 * NT];
#pragma unroll
    for (int j = 0; j < 2 * NT; ++j) bia[j] = p.bias ? __ldg(p.bias + t * 2 * NT + j) * (ACT == MGDT_ACT_SILU ? 0.5f : 1.f) : 0.f;

    pdl_wait();

    // ---- persistent CTA: strips blockIdx.x, + gridDim.x, ...; strip i + 1 is staged (zero-filling cp.async: rows y0 - 1 ..
    // y0 + TH, columns -1 .. TWp - 2) into the other buffer while strip i is computed
    auto stage = [&](int strip, uint32_t buf) {
        const int n = strip / p.tiles_y, y0 = (strip - n * p.tiles_y) * CW_TH;
        for (int r = 0; r < CW_TH + 2; ++r) {
            const int iy = y0 - 1 + r;
            const bool rowok = iy >= 0 && iy < p.H;
            const __nv_bfloat16* srow = p.x + ((size_t)n * p.H + (rowok ? iy : 0)) * p.W * p.x_cs;
            const uint32_t drow = buf + r * TWp * PITCH;
            for (int i = tid; i < TWp * CPP; i += CW_TH * 32) {
                const int cx = i / CPP, ch = i - cx * CPP, ix = cx - 1;
                const bool ok = rowok && ix >= 0 && ix < p.W;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(drow + cx * PITCH + ch * 16),
                             "l"(srow + (size_t)(ok ? ix : 0) * p.x_cs + ch * 8), "r"(ok ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (tid < CW_SLACK * CPP) {  // slack pixels of both buffers: read by the masked second tile of a pair, must be finite
        *reinterpret_cast<uint4*>(smem + ((CW_TH + 2) * TWp + tid / CPP) * PITCH + (tid % CPP) * 16) = make_uint4(0, 0, 0, 0);
        if (p.nbuf > 1) *reinterpret_cast<uint4*>(smem + bufsz + ((CW_TH + 2) * TWp + tid / CPP) * PITCH + (tid % CPP) * 16) = make_uint4(0, 0, 0, 0);
    }
    stage(blockIdx.x, sbase);


    // ldmatrix lane addresses.  A: lane l -> matrix m = l >> 3 (m & 1: pixels 8-15, m >> 1: second half of the K step),
    // row i = l & 7.  Cin >= 16: the second half is channels + 8 of the same tap; Cin = 8: it is the next tap.
    const int li = lane & 7, lm = lane >> 3;
    const uint32_t arow = (warp * TWp + li + 8 * (lm & 1)) * PITCH + (CIN == 8 ? 0 : 16 * (lm >> 1));   // + buffer base
    uint32_t aoff[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        const int tap = CIN == 8 ? min(2 * s + (lm >> 1), 8) : s / C16;
        aoff[s] = ((tap / 3) * TWp + tap % 3) * PITCH + (CIN == 8 ? 0 : (s % C16) * 32);
    }
    // B (32 -> 32): matrices {n-tile 2 jj, k lo}, {2 jj, k hi}, {2 jj + 1, k lo}, {2 jj + 1, k hi}
    const uint32_t brow = wbase + ((lm >> 1) * 8 + li) * WP + 16 * (lm & 1);

    int it = 0;
    for (int strip = blockIdx.x; strip < p.strips; strip += gridDim.x, ++it) {
    const uint32_t buf = sbase + (it & 1) * bufsz;
    if (strip + (int)gridDim.x < p.strips) {
        stage(strip + gridDim.x, sbase + ((it + 1) & 1) * bufsz);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int n = strip / p.tiles_y, oy = (strip - n * p.tiles_y) * CW_TH + warp;
    if (oy < p.H) {
    const size_t pix0 = ((size_t)n * p.H + oy) * p.W + g;
    __nv_bfloat16* dst = p.y + pix0 * p.y_cs + t * 2 * NT;
    const __nv_bfloat16* rsp = p.res ? p.res + pix0 * p.res_cs + t * 2 * NT : nullptr;
    const size_t dstep = (size_t)16 * MT * p.y_cs, rstep = (size_t)16 * MT * p.res_cs;
    uint32_t a0 = buf + arow;
    for (int xt = 0; xt < p.TX; xt += MT, a0 += 16 * MT * PITCH, dst += dstep, rsp += (p.res ? rstep : 0)) {
        Vec rs[MT][2];
        if (p.res) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if ((xt + m) * 16 + g + 8 * h < p.W) rs[m][h] = *reinterpret_cast<const Vec*>(rsp + (size_t)(16 * m + 8 * h) * p.res_cs);
        }
        float acc[MT][NT][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = acc[m][j][2] = acc[m][j][3] = 0.f;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            uint32_t a[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) ldsm4(a[m], a0 + aoff[s] + m * 16 * PITCH);
            if (B_REGS) {
#pragma unroll
                for (int j = 0; j < NT; ++j) mma_bf16_16816_(acc[0][j], a[0], bf[B_REGS ? s : 0][B_REGS ? j : 0][0], bf[B_REGS ? s : 0][B_REGS ? j : 0][1]);
            } else {
#pragma unroll
                for (int jj = 0; jj < NT / 2; ++jj) {
                    uint32_t b[4];
                    ldsm4(b, brow + jj * 16 * WP + s * 32);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        mma_bf16_16816_(acc[m][2 * jj], a[m], b[0], b[1]);
                        mma_bf16_16816_(acc[m][2 * jj + 1], a[m], b[2], b[3]);
                    }
                }
            }
        }
        // ---- epilogue: channels t * 2 NT .. of pixels x = xt * 16 + g (acc[.][0..1]) and x + 8 (acc[.][2..3])
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if ((xt + m) * 16 + g + 8 * h >= p.W) continue;
                alignas(16) uint32_t o[NT];
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(&rs[m][h]);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    float v0, v1;
                    if (ACT == MGDT_ACT_SILU) {      // h + h tanh(h), h = (acc + bias) / 2
                        const float h0 = fmaf(acc[m][j][2 * h], 0.5f, bia[2 * j]), h1 = fmaf(acc[m][j][2 * h + 1], 0.5f, bia[2 * j + 1]);
                        v0 = fmaf(h0, tanh_fast(h0), h0); v1 = fmaf(h1, tanh_fast(h1), h1);
                    } else {
                        v0 = act_fast_rt(acc[m][j][2 * h] + bia[2 * j], p.act); v1 = act_fast_rt(acc[m][j][2 * h + 1] + bia[2 * j + 1], p.act);
                    }
                    if (p.res) { v0 += bf_lo(rw[j]); v1 += bf_hi(rw[j]); }
                    o[j] = pack_bf2w(v0, v1);
                }
                *reinterpret_cast<Vec*>(dst + (size_t)(16 * m + 8 * h) * p.y_cs) = *reinterpret_cast<const Vec*>(o);
            }
    }
    }
    __syncthreads();   // every warp is done with this buffer before the next iteration stages strip i + 2 into it
    }
}

static size_t cw_smem(int Cin, int Cout, int W, int nbuf) {
    const int TX = (W + 15) / 16, TWp = TX * 16 + 4;
    const int pitch = Cin == 8 ? 16 : Cin * 2 + 16, ks = Cin == 8 ? 5 : 9 * (Cin / 16), nt = Cout / 8;
    return (size_t)nbuf * ((CW_TH + 2) * TWp + CW_SLACK) * pitch + (ks * nt <= 18 ? 0 : (size_t)Cout * (9 * Cin * 2 + 16));
}

template <int CIN, int COUT>
static int cw_launch(const CwP& p_in, cudaStream_t s) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    // one CTA per strip (default), or -- "conv3x3_warp_spc" > 1 -- persistent CTAs that walk >= spc strips with the next
    // strip's staging overlapped (double buffer); measured on B200: no gain (the second buffer costs occupancy)
    CwP p = p_in;
    unsigned nblk = (unsigned)p.strips;
    p.nbuf = 1;
    if (g_conv3x3_warp_spc > 1) {
        const size_t smem2 = cw_smem(CIN, COUT, p.W, 2);
        const int fit = (int)std::max<size_t>(1, std::min<size_t>(12, (200 * 1024) / (smem2 + 1024)));
        const int per_sm = std::max(1, std::min(fit, p.strips / (g_conv3x3_warp_spc * sms)));
        if (sms * per_sm < p.strips) { nblk = (unsigned)(sms * per_sm); p.nbuf = 2; }
    }
    const size_t smem = cw_smem(CIN, COUT, p.W, p.nbuf);
    const dim3 grid(nblk), block(CW_TH * 32);
    if (p.act == MGDT_ACT_SILU) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv3x3_warp_kernel<CIN, COUT, MGDT_ACT_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_k(conv3x3_warp_kernel<CIN, COUT, MGDT_ACT_SILU>, grid, block, smem, s, p);
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute(conv3x3_warp_kernel<CIN, COUT, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_k(conv3x3_warp_kernel<CIN, COUT, -1>, grid, block, smem, s, p);
    }
    MGDT_LAUNCH_CHECK("conv3x3_warp");
    return 0;
}

bool conv3x3_warp_supported(const mgdt_conv_args* a) {
    if (!g_conv3x3_warp || a->impl != 0 || a->dtype != MGDT_BF16) return false;
    if (a->kh != 3 || a->kw != 3 || a->stride != 1 || a->pad != 1) return false;
    // measured in the model at B = 32 (us, this kernel / tcgen05): 8 -> 8 @160^2 17.9 / 28.4, 16 -> 16 @80^2 14.8 / 18.4,
    // 32 -> 32 @40^2 15.4 / 14.4, 32 -> 32 @80^2 26.4 / 21.2.  32 -> 32 is slower per launch but the whole step is not (three
    // batches in flight: value +-0, end-to-end +0.9 % in three alternating A/B pairs) -- a 4-warp CTA grid shares the SMs
    // with the other streams' kernels, the persistent 148-CTA tcgen05 kernel does not -- so it is on by default too
    if (!(a->Cin == 8 || a->Cin == 16 || (a->Cin == 32 && g_conv3x3_warp >= 2)) || a->Cout != a->Cin) return false;
    if (a->pre_add || a->in_scale || a->pix_scale || a->in_relu || a->stat_acc || a->w_per_image || a->act_cols) return false;
    if ((a->x_cs & 7) || ((uintptr_t)a->x & 15) || ((uintptr_t)a->w & 15)) return false;                // 16-byte pixel / weight chunks
    const int vb = a->Cout / 2;                                                                            // bytes per thread store: 4 / 8 / 16
    if (((uintptr_t)a->y & (vb - 1)) || (a->y_cs * 2) % vb) return false;
    if (a->residual && (((uintptr_t)a->residual & (vb - 1)) || (a->res_cs * 2) % vb)) return false;
    return cw_smem(a->Cin, a->Cout, a->W, 2) <= 160 * 1024 && (long long)a->N * cdiv(a->H, CW_TH) < (1LL << 31);
}

int conv3x3_warp(const mgdt_conv_args* a, cudaStream_t s) {
    CwP p;
    p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w; p.res = (const __nv_bfloat16*)a->residual;
    p.bias = a->bias; p.y = (__nv_bfloat16*)a->y;
    p.N = a->N; p.H = a->H; p.W = a->W; p.x_cs = a->x_cs; p.y_cs = a->y_cs; p.res_cs = a->res_cs; p.act = a->act;
    p.TX = (a->W + 15) / 16; p.TWp = p.TX * 16 + 4; p.tiles_y = cdiv(a->H, CW_TH); p.strips = a->N * p.tiles_y;
    switch (a->Cin) {
        case 8: return cw_launch<8, 8>(p, s);
        case 16: return cw_launch<16, 16>(p, s);
        default: return cw_launch<32, 32>(p, s);
    }
}

}  // namespace mgdt
