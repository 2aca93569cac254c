// Persistent, warp-specialised tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
// Same GEMM mapping and shared-memory layouts as the first, non-persistent version (kept as conv_umma_v1.cu.txt for its
// layout description; not compiled): pixels on M = 128,
// channel planes [Cin/8][parity][P][16 B] in the no-swizzle K-major UMMA layout, taps as shifted
// descriptors, stride 2 as four parity sub-images), restructured as ONE resident CTA per SM that
// loops over output tiles with three overlapped roles:
//
//   warps 0-7   producers  stage A (and, for K-sliced layers, the matching weight slice) of work item i+1
//               into a ring of S shared-memory stages: 16-byte LDG -> fused input transform -> STS,
//               fence.proxy.async, mbarrier arrive (full[s])
//   warp  16    MMA        one lane waits full[s], issues the tcgen05.mma chain of the item into one of two
//               TMEM accumulator buffers, tcgen05.commit -> empty[s]  (+ accfull[a] on the last K slice)
//   warps 8-15  epilogue   wait accfull[a]; tcgen05.ld 32x32b -> bias/act (SFU)/residual -> bf16 -> NHWC stores;
//               arrive accempty[a]
//
// so the global loads of tile i+1, the tensor-core work of tile i and the stores of tile i-1 overlap,
// and TMEM allocation, barrier setup and the (resident) weight image are paid once per SM instead of
// once per tile.  Layers whose whole K extent does not fit next to its weights are K-sliced (nks > 1):
// each work item is (tile, K slice) and carries its own weight slice through the ring.
#include "common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#ifndef MGDT_WAIT_HINT
#define MGDT_WAIT_HINT 100000   // mbarrier.try_wait suspend-time hint (ns)
#endif

namespace mgdt {

// 20 warps (five per SM sub-partition, so 96 registers per thread): producers | one MMA warp | epilogue; the split
// depends on the loader (struct Roles).
constexpr int U2_WARPS = 20;
constexpr int U2_THREADS = U2_WARPS * 32;
constexpr int U2_MAX_EPI_WARPS = 16;
constexpr int U2_MAX_SMEM = 226 * 1024;
constexpr int U2_MAX_STAGES = 4;
constexpr int U2_MLP = 8;          // 16-byte loads in flight per producer thread
constexpr int U2_MAX_MMA = 160;   // K=16 instructions per (slice, 128-row block) the descriptor table holds
constexpr int U2_TAIL = 128 + 2 * 8 * U2_MAX_MMA + 4 * 256 + 2 * 256 + U2_MAX_EPI_WARPS * 2048 + 512;  // barriers + TMEM slot, descriptor tables, bias[Nc], u8 LUT, epilogue staging (512-byte aligned)

struct FastDiv {  // exact n / d for 0 <= n < 2^31
    uint32_t mul, shr, d;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1u << l) < d) ++l;
    const uint32_t s = 31 + l;
    f.shr = s;
    f.mul = (uint32_t)((((unsigned long long)1 << s) + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    return (uint32_t)(((unsigned long long)n * f.mul) >> f.shr);
}

static int g_force_mb = 0;  // option "conv_mb": force the row blocks per tile (1 / 2 / 4) where the plan allows it (experiments)
static int g_ksplit = 0;   // option "conv_ksplit": K-split partial accumulators (see try_run2 / plan_t1); measured: no gain (the MMA phase is not bound by accumulator dependencies), so off

struct Plan2 {
    int mode;      // 0: 1x1 s1 (also DCN), 1: 3x3 s1, 2: 3x3 s2
    int planes, npar, taps;
    int Npad, Nc, nsplit;
    int PS, nks;   // planes per K slice, number of slices
    int nmma_s;    // K=16 instructions per (slice, 128-row block) = ceil(taps*PS / 2)
    int tap_par[9], tap_dy[9], tap_dx[9];
    bool ok;
};

struct Run2 {
    int MB, S, NACC, P, Wq, halo, pstride16, tiles_per_img, tmem_cols;
    int KS;        // independent partial accumulators per row block (K-split): consecutive MMAs of a row block rotate over them
    int per_img;   // mode 0 tiled per image (tiles never straddle images): needed for per-image weights
    long long tiles;
    unsigned a_bytes, w_slice_bytes, stage_bytes, wres_bytes, smem_total;
};

constexpr long long U2_STAGE_BUDGET = U2_MAX_SMEM - U2_TAIL - 6 * 1024;   // weights + ring (estimates run a little low)

static Plan2 make_plan2(int Cin, int Cout, int k, int stride) {
    Plan2 p{};
    p.ok = false;
    if (Cin % 8 != 0 || Cin < 8 || Cout < 1 || Cout > 1024) return p;
    if (k == 1 && stride == 1) p.mode = 0;
    else if (k == 3 && stride == 1) p.mode = 1;
    else if (k == 3 && stride == 2) p.mode = 2;
    else return p;
    p.planes = Cin / 8;
    p.npar = p.mode == 2 ? 4 : 1;
    p.taps = p.mode == 0 ? 1 : 9;
    p.Npad = (Cout + 15) / 16 * 16;
    int n = 0;
    if (p.mode == 0) {
        p.tap_par[0] = 0; p.tap_dy[0] = 0; p.tap_dx[0] = 0;
    } else if (p.mode == 1) {
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) { p.tap_par[n] = 0; p.tap_dy[n] = dy; p.tap_dx[n] = dx; ++n; }
    } else {  // ascending (parity plane, shift): chunk pairs that straddle taps get a positive LBO
        for (int par = 0; par < 4; ++par)
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx)
                    if (((dy & 1) * 2 + (dx & 1)) == par) { p.tap_par[n] = par; p.tap_dy[n] = dy; p.tap_dx[n] = dx; ++n; }
    }
    // columns per CTA (<= 256, dividing Npad) and K slicing: the largest plane count per slice whose weight
    // slice and (nominal) A slice stay under ~80 KB each with two ring stages under ~200 KB; whole-K layers
    // keep their weights resident instead.
    const int a_plane_est = p.mode == 0 ? 2048 : (p.mode == 1 ? 4096 : 10240);
    int best = 0, Nc = p.Npad;
    while (Nc > 256 && Nc % 32 == 0) Nc /= 2;
    if (Nc > 256 || p.Npad % Nc != 0) return p;
    for (; Nc >= 16 && !best; Nc = (Nc % 32 == 0) ? Nc / 2 : 0) {
        for (int ps = p.planes; ps >= 1; --ps) {
            if (p.planes % ps) continue;
            if (p.taps > 1 && (ps & 1) && !(ps == 1 && p.planes == 1)) continue;  // pairs would straddle taps backwards
            const long long nm = (p.taps * ps + 1) / 2;
            if (nm > U2_MAX_MMA) continue;
            const long long wbytes = nm * 2 * Nc * 16;
            const long long abytes = (long long)ps * a_plane_est;
            const bool whole = ps == p.planes && wbytes <= 100 * 1024 && abytes <= 100 * 1024 && wbytes + 2 * abytes <= U2_STAGE_BUDGET;
            // sliced layers: A slices of at most ~40 KB, so that three ring stages fit next to resident weights -- with two
            // stages the producer of item k+2 waits for the MMAs of item k and the loop serialises (measured on 384->96)
            const bool sliced = wbytes <= 80 * 1024 && abytes <= 40 * 1024 && 2 * (wbytes + abytes) <= U2_STAGE_BUDGET;
            if (whole || sliced) { best = ps; break; }
        }
        if (best) break;
    }
    if (!best) return p;
    p.Nc = Nc;
    p.nsplit = p.Npad / Nc;
    p.PS = best;
    p.nks = p.planes / best;
    p.nmma_s = (p.taps * p.PS + 1) / 2;
    p.ok = true;
    return p;
}

static bool try_run2(const Plan2& p, int MB, int N, int H, int W, int Ho, int Wo, Run2& r, bool per_img_w = false) {
    r.MB = MB;
    r.NACC = (2 * MB * p.Nc <= 512) ? 2 : 1;
    r.per_img = 0;
    if (MB * p.Nc > 512) return false;
    if (p.mode == 0 && per_img_w) {
        r.Wq = W; r.halo = 0; r.P = 128 * MB;
        r.per_img = 1;
        r.tiles_per_img = (H * W + 128 * MB - 1) / (128 * MB);
        r.tiles = (long long)r.tiles_per_img * N;
    } else if (p.mode == 0) {
        r.Wq = W; r.halo = 0; r.P = 128 * MB;
        r.tiles_per_img = 1;
        r.tiles = ((long long)N * H * W + 128 * MB - 1) / (128 * MB);
    } else if (p.mode == 1) {
        r.Wq = W + 2; r.halo = r.Wq + 1; r.P = 128 * MB + 2 * r.halo;
        r.tiles_per_img = (H * r.Wq + 128 * MB - 1) / (128 * MB);
        r.tiles = (long long)r.tiles_per_img * N;
    } else {
        r.Wq = Wo + 1; r.halo = 0; r.P = 128 * MB + r.Wq + 1;
        r.tiles_per_img = (Ho * r.Wq + 128 * MB - 1) / (128 * MB);
        r.tiles = (long long)r.tiles_per_img * N;
    }
    r.P = (r.P + 7) / 8 * 8 + 8;
    r.pstride16 = p.npar * r.P;
    if ((r.pstride16 & 1) == 0) r.pstride16 += 1;
    r.a_bytes = (unsigned)(((size_t)p.PS * r.pstride16 * 16 + 127) / 128 * 128);
    const unsigned wall = (unsigned)((size_t)p.nmma_s * 2 * p.Nc * 16);
    r.w_slice_bytes = p.nks > 1 ? wall : 0;
    r.wres_bytes = p.nks > 1 ? 0 : wall;
    // K-sliced layers whose complete weight image still fits next to two A stages keep it resident (copied once per
    // CTA) instead of carrying every slice through the ring with every tile (384->96: 74 KB re-read per 128-row tile)
    if (p.nks > 1 && (size_t)p.nks * wall <= 112 * 1024 &&
        (size_t)p.nks * wall + 2 * (size_t)r.a_bytes + U2_TAIL <= (size_t)U2_MAX_SMEM) {
        r.wres_bytes = (unsigned)((size_t)p.nks * wall);
        r.w_slice_bytes = 0;
    }
    if (per_img_w) {   // per-image weights change with the tile: every item carries its slice through the ring
        r.wres_bytes = 0;
        r.w_slice_bytes = wall;
    }
    r.stage_bytes = r.a_bytes + r.w_slice_bytes;
    int S = U2_MAX_STAGES;
    while (S >= 2 && (size_t)r.wres_bytes + (size_t)S * r.stage_bytes + U2_TAIL > (size_t)U2_MAX_SMEM) --S;
    if (S < 2) return false;
    r.S = S;
    r.smem_total = r.wres_bytes + S * r.stage_bytes + U2_TAIL;
    // The MMAs of a tile are issued by up to four warps (Roles::NMW: one thread issues only one tcgen05.mma per ~154
    // cycles), each owning whole accumulation chains.  A chain = (row block, partial accumulator): the K = 16 steps of a
    // row block rotate over KS partial accumulators (summed by the epilogue) so that even a one-row-block tile has four
    // independent chains, MB * KS >= 4.
    r.KS = 1;
    const long long nmma_total = (long long)p.nmma_s * p.nks;
    if (g_ksplit)
        for (int ks = 4; ks >= 2; ks >>= 1)
            if (MB * (ks / 2) < 4 && r.NACC * ks * MB * p.Nc <= 512 && ks * 2 <= nmma_total) { r.KS = ks; break; }
    int cols = 32;
    while (cols < r.NACC * r.KS * MB * p.Nc) cols <<= 1;
    r.tmem_cols = cols;
    return true;
}

static bool make_run2(const Plan2& p, int N, int H, int W, int Ho, int Wo, Run2& r, bool per_img_w = false) {
    const int mbs[3] = {4, 2, 1};
    bool found = false;
    if (g_force_mb) {
        Run2 t;
        if (try_run2(p, g_force_mb, N, H, W, Ho, Wo, t, per_img_w)) { r = t; return true; }
    }
    for (int i = 0; i < 3; ++i) {
        Run2 t;
        if (!try_run2(p, mbs[i], N, H, W, Ho, Wo, t, per_img_w)) continue;
        if (t.NACC < 2 && mbs[i] > 1) continue;           // keep two accumulator buffers when a smaller tile allows it
        r = t;
        found = true;
        // (with the uniform-datapath MMA issue a K = 16 step costs ~150 cycles + ~40 per row block, so four row blocks per
        // tile win as soon as most SMs get a tile: forced MB = 4 measured +2 % images/s over the old ">= 3 x 148 tiles" rule)
        if (t.tiles * p.nsplit >= 120) break;
    }
    if (!found) {
        Run2 t;
        if (try_run2(p, 1, N, H, W, Ho, Wo, t, per_img_w)) { r = t; found = true; }
    }
    return found;
}

// ---------------------------------------------------------------------------------- weight packing
// OHWI bf16 [Cout][k][k][Cin] -> [nsplit][nks][2*nmma_s chunks: (tap, plane-in-slice)][Nc][8], zero padded
template <typename S, typename D> __device__ __forceinline__ D cvt_w(S v);
template <> __device__ __forceinline__ __nv_bfloat16 cvt_w<__nv_bfloat16, __nv_bfloat16>(__nv_bfloat16 v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_w<float, __nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half cvt_w<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __half cvt_w<__nv_bfloat16, __half>(__nv_bfloat16 v) { return __float2half_rn(__bfloat162float(v)); }

// in_scale != nullptr: `nimg` images are packed back to back, image n with its input channels scaled by
// in_scale[n][ci] (per-image weights W diag(s_n): the GRN / TaskDecomposition input scale moved onto the weights).
template <typename S, typename D>
__global__ void umma2_pack_kernel(const S* __restrict__ w, D* __restrict__ out, Plan2 p, int Cin, int Cout, int k) {
    // NO pdl_trigger(): the convolution kernels copy their resident weights BEFORE griddepcontrol.wait (they are constants
    // of the layer), so the kernel that follows a pack in the stream must not be scheduled before this grid has completed
    // (without a trigger the dependent launch happens at grid completion).  Found as NaNs on the first forward of a model
    // (lazy pack -> conv back to back) once the conv's set-up got faster.
    pdl_wait();
    const int cps = p.nmma_s * 2;
    const long long total = (long long)p.nsplit * p.nks * cps * p.Nc * 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % 8);
        const int nl = (int)((i / 8) % p.Nc);
        const int chunk = (int)((i / (8LL * p.Nc)) % cps);
        const int ks = (int)((i / (8LL * p.Nc * cps)) % p.nks);
        const int ns = (int)(i / (8LL * p.Nc * cps * p.nks));
        const int co = ns * p.Nc + nl;
        D v = cvt_w<float, D>(0.f);
        if (chunk < p.taps * p.PS && co < Cout) {
            const int t = chunk / p.PS, plane = ks * p.PS + chunk % p.PS;
            v = cvt_w<S, D>(w[(((long long)co * k + p.tap_dy[t]) * k + p.tap_dx[t]) * Cin + plane * 8 + j]);
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // K-major, SWIZZLE_NONE shared-memory matrix descriptor, version 1 (Blackwell)
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
    // no release fence: the caller's outstanding global stores / reductions need not be performed first
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a protocol bug traps (CUDA error) instead of hanging the GPU
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        // suspend-time hint (ns): the waiting warp sleeps in hardware instead of polling, so the epilogue / MMA
        // warps do not steal issue slots from the producers while they wait
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)MGDT_WAIT_HINT) : "memory");
        if (ok) return;
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// run-time plan of the TMA-fed 1x1 kernel (conv_tma1x1.cuh)
struct T1 {
    int KB, nkb;          // channels per K block (= swizzle atom: 64 / 32 / 16), K blocks
    int kb_stage, nst, S; // K blocks per ring stage, stages per tile, ring depth
    int Nsub, nsub;       // columns per CTA, column sub-splits of a packed column block
    int NACC, KS, tmem_cols, w_ring, swz, ctas_per_sm, per_img, tiles_per_img, epiw;
    unsigned a_kb_bytes, a_stage_bytes, w_stage_bytes, stage_bytes, w_bytes, smem_total;
    long long tiles;
};

// run-time plan of the TMA-fed 3x3 stride-1 kernel (conv_tma3x3.cuh)
struct T3 {
    int MB, R, Wq;            // row blocks per tile, image rows per tile, padded row pitch W + 2
    int PB, pstride;          // positions one TMA box writes ((R + 2) * Wq; stride 2: (R + 1) * Wq per parity sub-plane), plane stride in positions
    int s2, npar, Ppar;       // stride-2 layer: four parity sub-planes of Ppar positions per plane (stride 1: npar = 1, Ppar = pstride)
    int Ho, Wo, xoff2;        // output size; stride 2: channel coordinate of the SECOND pixel of a column pair (= x_cs)
    int S, NACC, tmem_cols, ctas_per_sm, tiles_per_img, epiw;
    unsigned tx_bytes, w_copy_bytes, w_bytes, stage_bytes, smem_total;
    long long tiles;
};

struct P2 {
    const __nv_bfloat16 *x, *w, *pre_add, *pix_scale, *residual;
    const __nv_bfloat16* row_scale;   // 1x1 layers: the per-pixel input scale applied to the accumulator row instead (stride ps_cs)
    const float *bias, *in_scale;
    __nv_bfloat16* y;
    int N, H, W, Cin, Cout, Ho, Wo;
    int x_cs, y_cs, add_cs, ps_cs, res_cs, act, in_relu;
    int act_cols;                // activation on output channels < act_cols only (0 = all); TMA kernel
    int y_vec, res_vec;
    int w_f16;                   // weights packed as fp16 (B operand format F16), activations stay bf16
    float out_scale;             // accumulator scale applied with the bias (1/255 for the uint8 stem, else 1)
    Plan2 pl;
    Run2 rn;
    unsigned M_total;            // mode 0: N*H*W
    FastDiv d_ps, d_P, d_Wq, d_HW, d_tpi, d_W, d_cgs;
    const __nv_bfloat16 *dcn_off, *dcn_mask;   // DCNv2 staging: virtual 9*dcn_cin-channel input
    int off_cs, mask_cs, mask_logit, dcn_cin;
    int dcn_wide;                // DCN sampler: pixels are 32-byte aligned, so a corner's two channel groups come in ONE 256-bit load
    // fused preprocess + stem: 3x3 stride-2 conv read straight from the NCHW uint8 / float source image
    const void* stem_src;
    int stem_u8, stem_C, stem_H, stem_W;
    FastDiv d_Wo;
    // fused per-(n, c) statistics of the (bf16-rounded) output, accumulated by the epilogue with fp64 atomics into
    // st_acc[n][st_Q + st_sq][Cout]: st_Q = 0 none / 1 total / 5 total + the four adaptive_avg_pool2d(2) windows
    // (rows [0, st_h0e) | [st_h1b, Ho), columns [0, st_w0e) | [st_w1b, Wo)), st_sq = sum of squares as the last plane
    // The accumulators are replicated st_R times (copy = tile % st_R, stride st_rs doubles) so that CTAs working on
    // neighbouring tiles of one image do not serialise on the same L2 lines; st_tot = 0 skips the total plane when the
    // windows partition the image (even Ho, Wo: mgdt_stats_finish derives it from the four window sums).
    size_t w_img_elems;          // per-image weights: elements between consecutive images' packed weights (0 = shared)
    alignas(64) CUtensorMap xmap;   // TMA load of the activation as a 2D (channels, pixels) tensor (conv_tma1x1.cuh)
    T1 t1;
    T3 t3;
    alignas(64) CUtensorMap ymap;   // TMA store of 32-row x 32-channel output units (mode 0): 2D (channels, pixels) or, per-image tiles, 3D (channels, pixels of an image, image)
    int tma_store;
    int pair_ok;                 // paired 16-column epilogue units allowed (debug: MGDT_CONV_PAIR=0 turns them off)
    double* st_acc;
    int st_Q, st_sq, st_h0e, st_h1b, st_w0e, st_w1b, st_R, st_tot;
    long long st_rs;
    FastDiv d_oHW, d_oW;
    unsigned long long* trace;   // debug: per-CTA phase timestamps (mgdt_debug_set_trace), normally NULL
    // A-operand descriptor of every K = 16 instruction of a (slice, 128-row block), relative to the stage base (tile
    // independent, filled by the host: fill_adesc).  Read from the constant bank with a uniform index, so the MMA warp's
    // descriptor arithmetic stays in uniform registers (no LDS + R2UR chain in front of every UTCHMMA).
    unsigned long long adesc[U2_MAX_MMA];
};

static_assert(sizeof(P2) <= 4096, "kernel parameter space");

// tile-relative output row m -> output pixel index, or -1 for junk / out-of-range rows
__device__ __forceinline__ int out_pixel2(const P2& p, uint32_t tile, uint32_t m) {
    if (p.pl.mode == 0 && p.rn.per_img) {
        const uint32_t n = fdiv(tile, p.d_tpi);
        const uint32_t q = (tile - n * p.rn.tiles_per_img) * (128u * p.rn.MB) + m;
        return q < (uint32_t)(p.H * p.W) ? (int)(n * (uint32_t)(p.H * p.W) + q) : -1;
    }
    if (p.pl.mode == 0) {
        const uint32_t pix = tile * (128u * p.rn.MB) + m;
        return pix < p.M_total ? (int)pix : -1;
    }
    const uint32_t n = fdiv(tile, p.d_tpi);
    const uint32_t tt = tile - n * p.rn.tiles_per_img;
    if (p.pl.mode == 1) {
        const uint32_t q = p.rn.Wq + tt * 128u * p.rn.MB + m;
        const uint32_t hp = fdiv(q, p.d_Wq), wp = q - hp * p.rn.Wq;
        if (hp < 1 || hp > (uint32_t)p.H || wp < 1 || wp > (uint32_t)p.W) return -1;
        return (int)((n * p.H + (hp - 1)) * p.W + (wp - 1));
    }
    const uint32_t q = tt * 128u * p.rn.MB + m;
    const uint32_t ho = fdiv(q, p.d_Wq), wo = q - ho * p.rn.Wq;
    if (ho >= (uint32_t)p.Ho || wo >= (uint32_t)p.Wo) return -1;
    return (int)((n * p.Ho + ho) * p.Wo + wo);
}

// DCNv2: chunk = 8 channels (cg) of tap `tap` at output pixel g: mask * bilinear(x, p0 + offset)
__device__ __forceinline__ uint4 dcn_chunk(const P2& p, uint32_t g, int plane) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (g < p.M_total) {
        const int tap = (int)fdiv((uint32_t)plane, p.d_cgs), cg = plane - tap * (p.dcn_cin / 8);
        const uint32_t n = fdiv(g, p.d_HW);
        const uint32_t rem = g - n * (uint32_t)(p.H * p.W);
        const int hq = (int)fdiv(rem, p.d_W), wq = (int)rem - hq * p.W;
        const __nv_bfloat16* ofp = p.dcn_off + (size_t)g * p.off_cs + 2 * tap;
        const float dy = __bfloat162float(ofp[0]), dx = __bfloat162float(ofp[1]);
        float m = __bfloat162float(p.dcn_mask[(size_t)g * p.mask_cs + tap]);
        if (p.mask_logit) m = sigmoidf_(m);
        const float py = (float)(hq + tap / 3 - 1) + dy, px = (float)(wq + tap % 3 - 1) + dx;
        if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
            const int y0 = (int)floorf(py), x0 = (int)floorf(px);
            const float ly = py - (float)y0, lx = px - (float)x0;
            const float hy = 1.f - ly, hx = 1.f - lx;
            const __nv_bfloat16* xn = p.x + (size_t)n * p.H * p.W * p.x_cs + cg * 8;
            const float wgt[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
            const int yy[4] = {y0, y0, y0 + 1, y0 + 1}, xx[4] = {x0, x0 + 1, x0, x0 + 1};
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                if (yy[c4] >= 0 && yy[c4] <= p.H - 1 && xx[c4] >= 0 && xx[c4] <= p.W - 1) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(xn + (size_t)(yy[c4] * p.W + xx[c4]) * p.x_cs));
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 t = __bfloat1622float2(h[j]);
                        f[2 * j] = fmaf(wgt[c4], t.x, f[2 * j]);
                        f[2 * j + 1] = fmaf(wgt[c4], t.y, f[2 * j + 1]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] *= m;
        }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return o;
}

// Stem: K index k = (dy*3 + dx)*C + c of the 3x3 stride-2 pad-1 window, zero beyond 9*C; value = src/255 for
// uint8 (BasePredictor.preprocess, predictor.py:127-129), rounded to bf16 exactly as the unfused path does.
__device__ __forceinline__ uint4 stem_chunk(const P2& p, uint32_t g, int plane) {
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (g < p.M_total) {
        const uint32_t n = fdiv(g, p.d_HW);                       // d_HW = Ho*Wo here
        const uint32_t rem = g - n * (uint32_t)(p.H * p.W);
        const int ho = (int)fdiv(rem, p.d_Wo), wo = (int)rem - ho * p.W;
        const int C = p.stem_C, KK = 9 * C;
        const size_t plane_sz = (size_t)p.stem_H * p.stem_W;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = plane * 8 + j;
            if (k < KK) {
                const int tap = k / C, c = k - tap * C;
                const int hi = 2 * ho + tap / 3 - 1, wi = 2 * wo + tap % 3 - 1;
                if (hi >= 0 && hi < p.stem_H && wi >= 0 && wi < p.stem_W) {
                    const size_t idx = ((size_t)n * C + c) * plane_sz + (size_t)hi * p.stem_W + wi;
                    f[j] = p.stem_u8 ? (float)__ldg(reinterpret_cast<const uint8_t*>(p.stem_src) + idx) / 255.0f
                                     : __ldg(reinterpret_cast<const float*>(p.stem_src) + idx);
                }
            }
        }
    }
    uint4 o;
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return o;
}

// Debug timeline: trace[(cta * 64 + slot)] = globaltimer ns.  slots: 0 start, 1 setup done, 2 weights resident,
// 3 end; per tile t (< 6): 8+8t fill done (producer warp 0), 9+8t MMAs issued, 10+8t accumulators ready
// (epilogue warp 8), 11+8t epilogue done.
__device__ __forceinline__ void trace_mark(const P2& p, int slot) {
    if (p.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + slot] = t;
    }
}

// ---------------------------------------------------------------------------------- async copy helpers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    // 16-byte global -> shared copy that bypasses the register file; src_bytes = 0 zero-fills (halo / tail rows)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    // the mbarrier arrival fires once every cp.async this thread issued so far has landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

enum : int { LD_ASYNC = 0, LD_XFORM = 1, LD_DCN = 2, LD_STEM_U8 = 3, LD_STEM_GEN = 4 };

template <int LOADER, int SPLIT> struct Roles {
    // producer warps 0 .. NPW-1.  cp.async loaders: a warp sustains about one 512-byte LDGSTS per ~200 cycles in this
    // kernel (measured; tools/ubench/cpasync.cu gives 50-115 cycles for a bare loop), so the producer/epilogue split
    // is chosen per layer by the host cost model (SPLIT 0/1/2 = 3/16, 7/12, 11/8 producer/epilogue warps).  DCN
    // sampling and the stem's byte gather are instruction-bound and take most of the CTA.
    // MMA warps.  tools/ubench/umma_rate.cu: one thread issues a tcgen05.mma every ~150-300 cycles (the latency-bound
    // instruction sequence around each UTCHMMA), four threads together one per 40-75 cycles whatever N <= 64, the
    // layout or the alignment -- so the tensor core is NOT the limit of the small-N layers.  Spreading the chains of a
    // tile over 2 / 4 issuing warps (NMW = 4 with 8 / 4 producer warps) was measured in this kernel and did not help:
    // the 3x3 32->32 tile went from 2.9 to 4.0 us (trace_conv.py), 19.5k -> 17.7k images/s, so NMW stays 1.
    // SPLIT 3 (experiment, option conv_split = 3): 8 producer, 4 MMA-issuing, 8 epilogue warps.
    static constexpr int NMW = ((LOADER == LD_ASYNC || LOADER == LD_XFORM) && SPLIT == 3) ? 4 : 1;
    static constexpr int NPW = (LOADER == LD_ASYNC || LOADER == LD_XFORM) ? (SPLIT == 0 ? 3 : SPLIT == 1 ? 7 : SPLIT == 2 ? 11 : 8)
                               : LOADER == LD_DCN ? 15 : 11;
    static constexpr int MMAW = NPW;                          // first MMA warp
    static constexpr int EPI0 = NPW + NMW;                    // first epilogue warp (a multiple of 4: quadrant = warp % 4)
    static constexpr int NEW = U2_WARPS - EPI0;               // epilogue warps (16 or 12), NEW / 4 per lane quadrant
    static constexpr int NP = NPW * 32;                       // producer threads = arrival count of full[] / wready
};

// packed fp32 pair arithmetic (FADD2 / FMUL2 / FFMA2): halves the epilogue's ALU instruction count
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void up2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// Source pixel of staged position `pos` (parity `par`) of a tile, or -1 (zero fill); n = image of that pixel.
template <int MODE>
__device__ __forceinline__ int src_pixel(const P2& p, uint32_t tile, uint32_t tt, uint32_t n_img, uint32_t pos, uint32_t par,
                                         uint32_t& n) {
    n = n_img;
    if (MODE == 0 && p.rn.per_img) {
        const uint32_t q = tt * (128u * p.rn.MB) + pos;
        return q < (uint32_t)(p.H * p.W) ? (int)(n_img * (uint32_t)(p.H * p.W) + q) : -1;
    } else if (MODE == 0) {
        const uint32_t g = tile * (128u * p.rn.MB) + pos;
        if (g >= p.M_total) return -1;
        n = fdiv(g, p.d_HW);
        return (int)g;
    } else if (MODE == 1) {
        const int q = (int)(p.rn.Wq + tt * 128u * p.rn.MB + pos) - p.rn.halo;
        if (q < 0) return -1;
        const uint32_t hp = fdiv((uint32_t)q, p.d_Wq), wp = (uint32_t)q - hp * p.rn.Wq;
        if (hp < 1 || hp > (uint32_t)p.H || wp < 1 || wp > (uint32_t)p.W) return -1;
        return (int)((n * p.H + (hp - 1)) * p.W + (wp - 1));
    } else {
        const uint32_t q = tt * 128u * p.rn.MB + pos;
        const uint32_t r = fdiv(q, p.d_Wq), c = q - r * p.rn.Wq;
        const int hi = 2 * (int)r + (int)(par >> 1) - 1, wi = 2 * (int)c + (int)(par & 1) - 1;
        if (hi < 0 || hi >= p.H || wi < 0 || wi >= p.W) return -1;
        return (int)((n * p.H + hi) * p.W + wi);
    }
}

// The fused input transforms on one staged 16-byte chunk (8 channels of plane `plane`) of source pixel `pix`.
__device__ __forceinline__ uint4 xform_apply(const P2& p, uint4 v, int pix, uint32_t n, int plane, const float* sc = nullptr) {
    if (sc && p.in_scale && !p.pre_add && !p.pix_scale && !p.in_relu) {
        // scale-only transform (GRN, TaskDecomposition) with the scales in registers: bf16 pair -> two fp32 by shift /
        // mask, one packed FMUL2, one F2FP per pair
        uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned long long x2 = pk2(__uint_as_float(w[j] << 16), __uint_as_float(w[j] & 0xffff0000u));
            float lo, hi;
            up2(mul2(x2, pk2(sc[2 * j], sc[2 * j + 1])), lo, hi);
            const __nv_bfloat162 o = __floats2bfloat162_rn(lo, hi);
            w[j] = *reinterpret_cast<const uint32_t*>(&o);
        }
        return v;
    }
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    if (p.pre_add) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.pre_add + (size_t)pix * p.add_cs + plane * 8));
        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(ah[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
    }
    if (p.in_scale && sc) {   // the caller keeps this (image, plane)'s scales in registers
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= sc[j];
    } else if (p.in_scale) {
        const float4* s = reinterpret_cast<const float4*>(p.in_scale + (size_t)n * p.Cin + plane * 8);
        const float4 s0 = __ldg(s), s1 = __ldg(s + 1);
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w; f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
    }
    if (p.pix_scale) {
        const float s = __bfloat162float(p.pix_scale[(size_t)pix * p.ps_cs]);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= s;
    }
    if (p.in_relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return v;
}

// In-place transform of the chunks thread `ptid` copied into stage `sA` (zero-filled chunks stay zero, as in the
// reference where padding is applied after the producing op).
template <int MODE>
__device__ __forceinline__ void xform_stage(const P2& p, unsigned char* sA, uint32_t tile, uint32_t tt, uint32_t n_img,
                                            int plane0, uint32_t chunks, int ptid, int NP) {
    for (uint32_t e0 = ptid; e0 < chunks; e0 += NP * 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint32_t e = e0 + u * NP;
            if (e < chunks) {
                const uint32_t rest = fdiv(e, p.d_ps);
                const uint32_t pll = e - rest * p.pl.PS;
                uint32_t pos = rest, par = 0, n;
                if (MODE == 2) { par = fdiv(rest, p.d_P); pos = rest - par * p.rn.P; }
                const int pix = src_pixel<MODE>(p, tile, tt, n_img, pos, par, n);
                if (pix >= 0) {
                    uint4* q = reinterpret_cast<uint4*>(sA + (pll * p.rn.pstride16 + par * p.rn.P + pos) * 16u);
                    *q = xform_apply(p, *q, pix, n, plane0 + (int)pll);
                }
            }
        }
    }
}

// Same, for the mapping in which a thread always stages the same plane (positions pos0, pos0 + step, ...): the
// per-(image, channel) input scales of that plane live in registers and are reloaded only when the image changes.
template <int MODE>
__device__ __forceinline__ void xform_stage_fixed(const P2& p, unsigned char* sA, uint32_t tile, uint32_t tt, uint32_t n_img,
                                                  int plane, uint32_t pll, uint32_t pos0, uint32_t step, float* sc, uint32_t sc_n) {
    // sc[8] / sc_n: this plane's input scales of image sc_n, prefetched by the caller when the item was issued
    const uint32_t Pn = (uint32_t)p.rn.P;
    for (uint32_t pos = pos0; pos < Pn; pos += 4 * step) {   // four independent chunks per iteration
        int pix[4];
        uint32_t n[4];
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t ps = pos + u * step;
            pix[u] = ps < Pn ? src_pixel<MODE>(p, tile, tt, n_img, ps, 0, n[u]) : -1;
            if (pix[u] >= 0) v[u] = *reinterpret_cast<const uint4*>(sA + (pll * p.rn.pstride16 + ps) * 16u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (pix[u] < 0) continue;
            if (p.in_scale && n[u] != sc_n) {
                const float4* s = reinterpret_cast<const float4*>(p.in_scale + (size_t)n[u] * p.Cin + plane * 8);
                const float4 s0 = __ldg(s), s1 = __ldg(s + 1);
                sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                sc_n = n[u];
            }
            *reinterpret_cast<uint4*>(sA + (pll * p.rn.pstride16 + pos + u * step) * 16u) = xform_apply(p, v[u], pix[u], n[u], plane, sc);
        }
    }
}

// Epilogue arithmetic on NV accumulator columns of one output row: bias, activation, residual, bf16 pack.
// opixB >= -1 with pairB: columns 16..31 belong to a SECOND output row (paired 16-column units, see the epilogue).
template <int NV>
__device__ __forceinline__ void epi_math(const P2& p, const uint32_t* r, const float* sBias, int cbase, int co0, int opix,
                                         uint32_t* packed, bool pairB = false, int opixB = -1) {
    float v[NV];
    unsigned long long a[NV / 2];
    const unsigned long long half2 = pk2(0.5f, 0.5f);
    // acc * scale + bias: scale = out_scale (1 except for the uint8 stem) times, for 1x1 layers with a per-PIXEL input scale,
    // that pixel's scale (W (s x) = s (W x): the scale of cv3(cls_feat * cls_prob) is applied to the accumulator row)
    float rs = p.out_scale, rsB = p.out_scale;
    if (p.row_scale) {
        if (opix >= 0) rs *= __bfloat162float(p.row_scale[(size_t)opix * p.ps_cs]);
        if (pairB && opixB >= 0) rsB *= __bfloat162float(p.row_scale[(size_t)opixB * p.ps_cs]);
    }
    const unsigned long long sc2 = pk2(rs, rs), sc2B = pk2(rsB, rsB);
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(sBias + cbase);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
        const ulonglong2 b = b2[j];
        const unsigned long long sc = (pairB && 4 * j >= 16) ? sc2B : sc2;
        a[2 * j] = fma2(pk2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])), sc, b.x);
        a[2 * j + 1] = fma2(pk2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), sc, b.y);
    }
    if (p.act == MGDT_ACT_SILU || p.act == MGDT_ACT_SIGMOID) {
        // h = v/2, t = tanh(h) on the SFU: silu = h + h*t, sigmoid = 0.5 + 0.5*t
        const bool silu = p.act == MGDT_ACT_SILU;
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            const unsigned long long h = mul2(a[j], half2);
            float h0, h1;
            up2(h, h0, h1);
            const unsigned long long t = pk2(tanh_fast(h0), tanh_fast(h1));
            a[j] = silu ? fma2(h, t, h) : fma2(t, half2, half2);
        }
    }
    if (p.act == MGDT_ACT_GELU) {
        // packed form of act_fast<GELU>: erf(z) ~= tanh(z * (A + B z^2 + C z^4)), gelu = h + h * t with h = x / 2
        const unsigned long long kz = pk2(0.70710678118654752440f, 0.70710678118654752440f);
        const unsigned long long ka = pk2(MGDT_ERF_A, MGDT_ERF_A), kb = pk2(MGDT_ERF_B, MGDT_ERF_B), kc = pk2(MGDT_ERF_C, MGDT_ERF_C);
#pragma unroll
        for (int j = 0; j < NV / 2; ++j) {
            float z0, z1;
            up2(mul2(a[j], kz), z0, z1);
            const unsigned long long z = pk2(fminf(fmaxf(z0, -5.0f), 5.0f), fminf(fmaxf(z1, -5.0f), 5.0f));
            const unsigned long long u = mul2(z, z);
            float w0, w1;
            up2(mul2(z, fma2(fma2(kc, u, kb), u, ka)), w0, w1);
            const unsigned long long t = pk2(tanh_fast(w0), tanh_fast(w1));
            const unsigned long long h = mul2(a[j], half2);
            a[j] = fma2(h, t, h);
        }
    }
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) up2(a[j], v[2 * j], v[2 * j + 1]);
    switch (p.act) {
#define MGDT_ACT_CASE(A) case A: _Pragma("unroll") for (int j = 0; j < NV; ++j) v[j] = act_fast<A>(v[j]); break;
        MGDT_ACT_CASE(MGDT_ACT_RELU)
        MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
#undef MGDT_ACT_CASE
        default: break;
    }
    if (p.residual && (opix >= 0 || (pairB && opixB >= 0))) {
#pragma unroll
        for (int c8 = 0; c8 < NV; c8 += 8) {
            const bool second = pairB && c8 >= 16;            // paired unit: columns 16.. are channels 0.. of row opixB
            const int op = second ? opixB : opix, cc = second ? c8 - 16 : c8;
            if (op < 0 || co0 + cc >= p.Cout) continue;
            const __nv_bfloat16* rp = p.residual + (size_t)op * p.res_cs + co0 + cc;
            if (co0 + cc + 8 <= p.Cout && p.res_vec) {
                const uint4 ra = __ldg(reinterpret_cast<const uint4*>(rp));
                const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&ra);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(ah[j]); v[c8 + 2 * j] += t.x; v[c8 + 2 * j + 1] += t.y; }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) if (co0 + cc + j < p.Cout) v[c8 + j] += __bfloat162float(rp[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        packed[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
}

// Epilogue arithmetic of one unit (NV accumulator columns of one output row per lane), specialised on the activation so
// that the executed path is a straight run of packed fp32x2 instructions (FFMA2 / FMUL2 on adjacent register pairs,
// no repacking moves, no runtime switch inside the loop): SiLU is 2 FFMA2 + 2 MUFU.TANH + 1 F2FP per output PAIR.
//   h = acc * sc + hb   with (sc, hb) = (1/2, bias/2) for SiLU / sigmoid / GELU (the half of x the tanh forms need),
//                       (1, bias) otherwise; a per-pixel row scale (cv3(cls_feat * cls_prob)) multiplies sc
// sB holds the bias pre-multiplied accordingly (t1_bias_scale).
__host__ __device__ constexpr float t1_bias_scale(int act) {
    return (act == MGDT_ACT_SILU || act == MGDT_ACT_SIGMOID || act == MGDT_ACT_GELU) ? 0.5f : 1.0f;
}
// pairB (paired 16-column units of conv_umma2_kernel): columns 16..31 are channels 0..15 of a SECOND output row opixB.
template <int ACT, int NV>
__device__ __forceinline__ void epi_fast(const P2& p, const uint32_t* r, const float* sB, int cbase, int co0, int opix,
                                         uint32_t* packed, bool pairB = false, int opixB = -1) {
    float s = t1_bias_scale(ACT) * p.out_scale;
    if (p.row_scale && opix >= 0) s *= __bfloat162float(p.row_scale[(size_t)opix * p.ps_cs]);
    const float2 sc = make_float2(s, s);
    const float2* b2 = reinterpret_cast<const float2*>(sB + cbase);
    float2 y[NV / 2];
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
        const float2 acc = make_float2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        const float2 h = __ffma2_rn(acc, sc, b2[j]);
        if (ACT == MGDT_ACT_SILU) {
            const float2 t = make_float2(tanh_fast(h.x), tanh_fast(h.y));
            y[j] = __ffma2_rn(h, t, h);                                    // x * sigmoid(x) = h + h * tanh(h), h = x / 2
        } else if (ACT == MGDT_ACT_SIGMOID) {
            const float2 t = make_float2(tanh_fast(h.x), tanh_fast(h.y));
            y[j] = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
        } else if (ACT == MGDT_ACT_GELU) {
            // erf(z) ~= tanh(z * (A + B z^2 + C z^4)), z = x / sqrt(2) = h * sqrt(2), |z| clamped to 5 (common.cuh: act_fast)
            float2 z = __fmul2_rn(h, make_float2(1.41421356237309504880f, 1.41421356237309504880f));
            z.x = fminf(fmaxf(z.x, -5.0f), 5.0f);
            z.y = fminf(fmaxf(z.y, -5.0f), 5.0f);
            const float2 u = __fmul2_rn(z, z);
            const float2 q = __ffma2_rn(__ffma2_rn(make_float2(MGDT_ERF_C, MGDT_ERF_C), u, make_float2(MGDT_ERF_B, MGDT_ERF_B)), u,
                                        make_float2(MGDT_ERF_A, MGDT_ERF_A));
            const float2 w = __fmul2_rn(z, q);
            const float2 t = make_float2(tanh_fast(w.x), tanh_fast(w.y));
            y[j] = __ffma2_rn(h, t, h);
        } else if (ACT == MGDT_ACT_RELU) {
            y[j] = make_float2(fmaxf(h.x, 0.f), fmaxf(h.y, 0.f));
        } else if (ACT == MGDT_ACT_HSIGMOID) {
            y[j] = make_float2(__saturatef(fmaf(h.x, 1.0f / 6.0f, 0.5f)), __saturatef(fmaf(h.y, 1.0f / 6.0f, 0.5f)));
        } else {
            y[j] = h;
        }
    }
    if (p.residual && (opix >= 0 || (pairB && opixB >= 0))) {
#pragma unroll
        for (int c8 = 0; c8 < NV; c8 += 8) {
            const bool second = pairB && c8 >= 16;            // paired unit: columns 16.. are channels 0.. of row opixB
            const int op = second ? opixB : opix, cc = second ? c8 - 16 : c8;
            if (op < 0 || co0 + cc >= p.Cout) continue;
            const __nv_bfloat16* rp = p.residual + (size_t)op * p.res_cs + co0 + cc;
            if (co0 + cc + 8 <= p.Cout && p.res_vec) {
                const uint4 ra = __ldg(reinterpret_cast<const uint4*>(rp));
                const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&ra);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 tt = __bfloat1622float2(ah[j]);
                    y[c8 / 2 + j].x += tt.x;
                    y[c8 / 2 + j].y += tt.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (co0 + cc + j < p.Cout) {
                        const float add = __bfloat162float(rp[j]);
                        if (j & 1) y[(c8 + j) / 2].y += add; else y[(c8 + j) / 2].x += add;
                    }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV / 2; ++j) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(y[j].x, y[j].y);
        packed[j] = *reinterpret_cast<const uint32_t*>(&hh);
    }
}
template <int NV>
__device__ __forceinline__ void epi_fast_rt(const P2& p, const uint32_t* r, const float* sB, int cbase, int co0, int opix,
                                            uint32_t* packed, bool pairB = false, int opixB = -1, int act = -1) {
    switch (act < 0 ? p.act : act) {   // uniform per unit
        case MGDT_ACT_SILU: epi_fast<MGDT_ACT_SILU, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
        case MGDT_ACT_RELU: epi_fast<MGDT_ACT_RELU, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
        case MGDT_ACT_GELU: epi_fast<MGDT_ACT_GELU, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
        case MGDT_ACT_SIGMOID: epi_fast<MGDT_ACT_SIGMOID, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
        case MGDT_ACT_HSIGMOID: epi_fast<MGDT_ACT_HSIGMOID, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
        default: epi_fast<MGDT_ACT_NONE, NV>(p, r, sB, cbase, co0, opix, packed, pairB, opixB); break;
    }
}

// Fused output statistics: the calling epilogue warp has just staged a 32-row x 32-column bf16 unit in its swizzled
// shared-memory tile (row r at r*64, 16-byte chunk c at slot c ^ ((r >> 1) & 3)).  Lane l owns the column PAIR
// (2*(l & 15), +1) over the 16 rows r = 2*i + (l >> 4) (an even and an odd row per instruction: all 32 banks, no
// conflict), sums value and square with packed fp32x2 arithmetic, the two half-warps are combined by one shuffle and
// lanes 0-15 add the result to the per-(image, plane, channel) fp64 accumulators: total, sum of squares and the 2x2
// adaptive-pool windows the rows belong to.  skey = (image << 4) | window mask of the lane's own OUTPUT row
// (0xffffffff for junk rows); one masked pass per distinct key among the 32 rows (one in the common case, two or
// three when the group straddles a window or image boundary).  fp64 atomics make the result independent of the
// arrival order to ~1e-16, i.e. reproducible after the cast to fp32.
__device__ __forceinline__ void epi_stats(double* acc, int Q, int sq, int want_tot, int C, uint32_t stg32, uint32_t skey, int lane,
                                          int nv, int co0, int Cout) {
    const uint32_t full = 0xffffffffu, INVALID = 0xffffffffu;
    const uint32_t cp = (uint32_t)lane & 15u, half = (uint32_t)lane >> 4;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // row 2i + half: (row >> 1) & 3 == i & 3 for both halves
        const uint32_t a = stg32 + (uint32_t)(2 * i) * 64u + half * 64u + (((cp >> 2) ^ ((uint32_t)i & 3u)) << 4) + (cp & 3u) * 4u;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(a));
    }
    const int K = Q + sq;
    const int c = co0 + 2 * (int)cp;
    const bool col0 = half == 0 && 2 * (int)cp < nv && c < Cout, col1 = half == 0 && 2 * (int)cp + 1 < nv && c + 1 < Cout;
    uint32_t rem = __ballot_sync(full, skey != INVALID);
    while (rem) {
        const uint32_t k = __shfl_sync(full, skey, __ffs((int)rem) - 1);
        const uint32_t m = __ballot_sync(full, skey == k);
        rem &= ~m;
        const uint32_t mh = m >> half;
        unsigned long long sa[4], qa[4];   // four independent chains (fixed combination order)
#pragma unroll
        for (int u = 0; u < 4; ++u) sa[u] = qa[u] = pk2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t t = (m == full || (mh & (1u << (2 * i)))) ? w[i] : 0u;
            const unsigned long long v2 = pk2(__uint_as_float(t << 16), __uint_as_float(t & 0xffff0000u));
            sa[i & 3] = add2(v2, sa[i & 3]);
            qa[i & 3] = fma2(v2, v2, qa[i & 3]);
        }
        float s0, s1, q0, q1;
        up2(add2(add2(sa[0], sa[1]), add2(sa[2], sa[3])), s0, s1);
        up2(add2(add2(qa[0], qa[1]), add2(qa[2], qa[3])), q0, q1);
        s0 += __shfl_xor_sync(full, s0, 16); s1 += __shfl_xor_sync(full, s1, 16);
        q0 += __shfl_xor_sync(full, q0, 16); q1 += __shfl_xor_sync(full, q1, 16);
        if (col0) {
            double* base = acc + ((size_t)(k >> 4) * K) * C + c;
            if (want_tot) { atomicAdd(base, (double)s0); if (col1) atomicAdd(base + 1, (double)s1); }
            if (Q == 5) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k & (1u << j)) {
                        atomicAdd(base + (size_t)(1 + j) * C, (double)s0);
                        if (col1) atomicAdd(base + (size_t)(1 + j) * C + 1, (double)s1);
                    }
            }
            if (sq) { atomicAdd(base + (size_t)Q * C, (double)q0); if (col1) atomicAdd(base + (size_t)Q * C + 1, (double)q1); }
        }
    }
}

// r16[0..15] += the 16 accumulator columns at taddr (this lane's row): the partial accumulators of a K-split tile
__device__ __forceinline__ void tmem_ld_add16(uint32_t taddr, uint32_t* r16) {
    uint32_t t[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
          "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) r16[j] = __float_as_uint(__uint_as_float(r16[j]) + __uint_as_float(t[j]));
}

template <int MODE, int LOADER, int SPLIT, int STATS>
__global__ void __launch_bounds__(U2_THREADS, 1) conv_umma2_kernel(const __grid_constant__ P2 p) {
    pdl_trigger();
    extern __shared__ __align__(128) unsigned char smem[];
    const Plan2& pl = p.pl;
    const Run2& rn = p.rn;
    const int tid = threadIdx.x, lane = tid & 31;
    // warp index through a shuffle: provably warp-uniform, so the role dispatch below is a uniform branch and ptxas keeps
    // the MMA warp's loop state and descriptor arithmetic on the uniform datapath
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int ns = blockIdx.y;
    constexpr int NPW = Roles<LOADER, SPLIT>::NPW, NEW = Roles<LOADER, SPLIT>::NEW, NP = Roles<LOADER, SPLIT>::NP;
    constexpr int MMAW = Roles<LOADER, SPLIT>::MMAW, EPI0 = Roles<LOADER, SPLIT>::EPI0, NMW = Roles<LOADER, SPLIT>::NMW;
    static_assert(EPI0 % 4 == 0, "epilogue warps must start at a multiple of four");

    if (tid == 0) trace_mark(p, 0);
    unsigned char* sWres = smem;
    unsigned char* sStage = smem + rn.wres_bytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + rn.wres_bytes + (size_t)rn.S * rn.stage_bytes);
    // barrier slots: full[S] | empty[S] | accfull[2] | accempty[2] | wready
    const uint32_t bar0 = s_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (U2_MAX_STAGES + s); };
    auto ACCFULL = [&](int a) { return bar0 + 8u * (2 * U2_MAX_STAGES + a); };
    auto ACCEMPTY = [&](int a) { return bar0 + 8u * (2 * U2_MAX_STAGES + 2 + a); };
    const uint32_t WREADY = bar0 + 8u * (2 * U2_MAX_STAGES + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * U2_MAX_STAGES + 5);
    // per-instruction descriptor templates (tile independent): start offsets relative to the stage / weight base
    ulonglong2* desc_t = reinterpret_cast<ulonglong2*>(bars + 16);         // {A descriptor, B descriptor} per MMA
    float* sBias = reinterpret_cast<float*>(desc_t + U2_MAX_MMA);
    unsigned short* sLut = reinterpret_cast<unsigned short*>(sBias + 256);   // bf16(u / 255), u = 0..255
    unsigned char* sOut = reinterpret_cast<unsigned char*>(sLut + 256);      // epilogue staging: 2 KB per epilogue warp
    if (LOADER == LD_STEM_U8) {
        for (int i = tid; i < 256; i += U2_THREADS) sLut[i] = __bfloat16_as_ushort(__float2bfloat16_rn((float)i / 255.0f));
    }
    for (int i = tid; i < (pl.Nc == 16 ? 32 : pl.Nc); i += U2_THREADS) {   // Nc = 16: duplicated for the paired units
        const int co = ns * pl.Nc + (pl.Nc == 16 ? (i & 15) : i);
        sBias[i] = (p.bias && co < p.Cout) ? p.bias[co] : 0.f;
    }
    if (warp == MMAW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                     "r"((uint32_t)rn.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < rn.S; ++s) { mbar_init(FULL(s), NP); mbar_init(EMPTY(s), NMW); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACCFULL(a), NMW); mbar_init(ACCEMPTY(a), NEW); }
        mbar_init(WREADY, NP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) trace_mark(p, 1);
    // Everything above touched only shared memory, TMEM and this layer's constant parameters.  The activations may
    // still be being written by the previous kernel of the stream: each role calls pdl_wait() before its first access
    // to them (the producers after they have started the copy of the resident weights).

    const uint32_t tiles = (uint32_t)rn.tiles;
    const int nks = pl.nks;
    const size_t w_slice_elems = (size_t)pl.nmma_s * 2 * pl.Nc * 8;  // bf16 elements of one (ns, ks) weight slice

    if (warp < NPW) {
        // =============================================================== producers
        const int ptid = tid;  // 0..NP-1
        if (rn.wres_bytes) {   // resident weights: the whole (ns) image, all K slices
            const uint4* src = reinterpret_cast<const uint4*>(p.w + (size_t)ns * nks * w_slice_elems);
            const uint32_t dst = s_u32(sWres);
            const int n16 = (int)(rn.wres_bytes / 16);
            for (int i = ptid; i < n16; i += NP) cp_async16(dst + 16u * i, src + i, 16u);
            cp_async_arrive_noinc(WREADY);
            if (tid == 0) trace_mark(p, 2);
        }
        pdl_wait();
        const uint32_t chunks = (uint32_t)(pl.PS * pl.npar * rn.P);
        // (the in-place transform revisits chunks through the generic mapping, which matches the fast path's except for
        // the stride-2 parity loop)
        // LD_XFORM (not stride 2): use a multiple of PS threads so that every thread owns one plane (its input scales stay
        // in registers across the in-place transform); the few leftover threads only arrive on the barriers
        const int NPe = (LOADER == LD_XFORM && MODE != 2 && pl.PS <= NP) ? (NP / pl.PS) * pl.PS : NP;
        const bool ps_divides = NPe % pl.PS == 0 && !(LOADER == LD_XFORM && MODE == 2);
        const uint32_t pstep = ps_divides ? (uint32_t)(NPe / pl.PS) : 1u;
        const bool p_active = ptid < NPe;
        const uint32_t pos_fix = fdiv((uint32_t)ptid, p.d_ps), pll_fix = (uint32_t)ptid - pos_fix * pl.PS;
        uint32_t it = 0;
        int s = 0;
        uint32_t ph = 0;   // phase of the stage's current use (flips each time the ring wraps)
        float xf_sc[8], xf_sc_new[8];   // LD_XFORM: input scales of the pending item / of the item being issued
        uint32_t xf_scn = 0xffffffffu, xf_scn_new = 0xffffffffu;
        uint32_t stem_wa[9], stem_wb[9];   // LD_STEM_U8: prefetched source words
        bool xf_valid = false;   // LD_XFORM: the item whose raw copies are in flight and still to be transformed
        int xf_s = 0, xf_plane0 = 0;
        uint32_t xf_tile = 0, xf_tt = 0, xf_nimg = 0;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            uint32_t n_img = 0, tt = 0;
            if (MODE != 0 || rn.per_img) { n_img = fdiv(tile, p.d_tpi); tt = tile - n_img * rn.tiles_per_img; }
            for (int ks = 0; ks < nks; ++ks, ++it) {
                mbar_wait(EMPTY(s), ph ^ 1);
                unsigned char* sA = sStage + (size_t)s * rn.stage_bytes;
                const uint32_t sA32 = s_u32(sA);
                if (rn.w_slice_bytes) {
                    const uint4* src = reinterpret_cast<const uint4*>(p.w + (size_t)n_img * p.w_img_elems + ((size_t)ns * nks + ks) * w_slice_elems);
                    const uint32_t dst = sA32 + rn.a_bytes;
                    const int n16 = (int)(rn.w_slice_bytes / 16);
                    for (int i = ptid; i < n16; i += NP) cp_async16(dst + 16u * i, src + i, 16u);
                }
                const int plane0 = ks * pl.PS;
                if (LOADER == LD_XFORM && p.in_scale && ps_divides && p_active) {
                    // prefetch this item's per-(image, channel) input scales; they are used one iteration later
                    uint32_t n0 = n_img;
                    if (MODE == 0) n0 = fdiv(min(tile * (128u * rn.MB) + pos_fix, p.M_total - 1u), p.d_HW);
                    const float4* sp = reinterpret_cast<const float4*>(p.in_scale + (size_t)n0 * p.Cin + (plane0 + (int)pll_fix) * 8);
                    const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                    xf_sc_new[0] = s0.x; xf_sc_new[1] = s0.y; xf_sc_new[2] = s0.z; xf_sc_new[3] = s0.w;
                    xf_sc_new[4] = s1.x; xf_sc_new[5] = s1.y; xf_sc_new[6] = s1.z; xf_sc_new[7] = s1.w;
                    xf_scn_new = n0;
                }
                if (LOADER == LD_ASYNC || LOADER == LD_XFORM) {
                    // transform-free input: every chunk is an asynchronous 16-byte copy (zero-filled outside the image);
                    // the thread never waits for its data, the stage's full barrier counts the copies' completion
                    if (!p_active) {
                        // spare thread of the fixed-plane mapping: nothing to copy, it only takes part in the barriers
                    } else if (ps_divides) {
                        // NP % PS == 0: this thread always stages the same plane and its positions advance by a constant
                        // step.  Four independent chunks per iteration (index arithmetic by multiply-shift division, no
                        // data-dependent branches) so the address chains overlap.
                        const __nv_bfloat16* xpl = p.x + (plane0 + pll_fix) * 8;
                        const uint32_t dpl = sA32 + (uint32_t)pll_fix * rn.pstride16 * 16u;
                        const uint32_t Pn = (uint32_t)rn.P;
                        if (MODE == 0) {
                            // linear pixel tiles, or (per-image weights) tiles that restart at every image
                            const uint32_t g0 = rn.per_img ? n_img * (uint32_t)(p.H * p.W) + tt * (128u * rn.MB) : tile * (128u * rn.MB);
                            const uint32_t gend = rn.per_img ? (n_img + 1u) * (uint32_t)(p.H * p.W) : p.M_total;
                            for (uint32_t pos0 = pos_fix; pos0 < Pn; pos0 += 4 * pstep) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const uint32_t pos = pos0 + u * pstep;
                                    const uint32_t g = g0 + pos;
                                    const bool ok = g < gend;
                                    if (pos < Pn) cp_async16(dpl + pos * 16u, ok ? xpl + (size_t)g * p.x_cs : p.x, ok ? 16u : 0u);
                                }
                            }
                        } else if (MODE == 1) {
                            // padded-linear index shifted by one padded row so that it is never negative:
                            // hp = padded row + 1, valid rows 2 .. H+1, valid columns 1 .. W
                            const uint32_t q1 = tt * 128u * rn.MB + (uint32_t)rn.Wq - 1u;
                            const int pixc = (int)(n_img * p.H * p.W) - 2 * p.W - 1;
                            for (uint32_t pos0 = pos_fix; pos0 < Pn; pos0 += 4 * pstep) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const uint32_t pos = pos0 + u * pstep;
                                    const uint32_t q = q1 + pos;
                                    const uint32_t hp = fdiv(q, p.d_Wq), wp = q - hp * rn.Wq;
                                    const bool ok = (hp - 2u) < (uint32_t)p.H && (wp - 1u) < (uint32_t)p.W;
                                    const int pix = (int)(hp * p.W + wp) + pixc;
                                    if (pos < Pn) cp_async16(dpl + pos * 16u, ok ? xpl + (size_t)pix * p.x_cs : p.x, ok ? 16u : 0u);
                                }
                            }
                        } else {
                            const uint32_t q0 = tt * 128u * rn.MB;
                            const int pixc = (int)(n_img * p.H * p.W);
                            for (uint32_t par = 0; par < 4; ++par) {
                                const int dh = (int)(par >> 1) - 1, dw = (int)(par & 1) - 1;
                                const uint32_t dpar = dpl + par * Pn * 16u;
                                for (uint32_t pos0 = pos_fix; pos0 < Pn; pos0 += 4 * pstep) {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const uint32_t pos = pos0 + u * pstep;
                                        const uint32_t q = q0 + pos;
                                        const uint32_t r = fdiv(q, p.d_Wq), c = q - r * rn.Wq;
                                        const int hi = 2 * (int)r + dh, wi = 2 * (int)c + dw;
                                        const bool ok = (uint32_t)hi < (uint32_t)p.H && (uint32_t)wi < (uint32_t)p.W;
                                        const int pix = hi * p.W + wi + pixc;
                                        if (pos < Pn) cp_async16(dpar + pos * 16u, ok ? xpl + (size_t)pix * p.x_cs : p.x, ok ? 16u : 0u);
                                    }
                                }
                            }
                        }
                    } else
                    for (uint32_t e = ptid; e < chunks; e += NPe) {
                        const uint32_t rest = fdiv(e, p.d_ps);
                        const uint32_t pll = e - rest * pl.PS;
                        uint32_t pos = rest, par = 0, n;
                        if (MODE == 2) { par = fdiv(rest, p.d_P); pos = rest - par * rn.P; }
                        const int pix = src_pixel<MODE>(p, tile, tt, n_img, pos, par, n);
                        const __nv_bfloat16* src = pix >= 0 ? p.x + (size_t)pix * p.x_cs + (plane0 + pll) * 8 : p.x;
                        cp_async16(sA32 + (pll * rn.pstride16 + par * rn.P + pos) * 16u, src, pix >= 0 ? 16u : 0u);
                    }
                    if (LOADER == LD_ASYNC) cp_async_arrive_noinc(FULL(s));
                    else {
                        // fused input transforms: the raw copies of THIS item stay in flight while the thread transforms,
                        // in place, the chunks it copied for the PREVIOUS item (same thread -> chunk mapping, so no
                        // cross-thread synchronisation), then publishes that stage
                        asm volatile("cp.async.commit_group;" ::: "memory");
                        if (xf_valid) {
                            asm volatile("cp.async.wait_group 1;" ::: "memory");
                            if (tid == 0 && it < 6) trace_mark(p, 12 + 8 * (int)it);
                            if (p_active) {
                                if (ps_divides) xform_stage_fixed<MODE>(p, sStage + (size_t)xf_s * rn.stage_bytes, xf_tile, xf_tt, xf_nimg,
                                                                        xf_plane0 + (int)pll_fix, pll_fix, pos_fix, pstep, xf_sc, xf_scn);
                                else xform_stage<MODE>(p, sStage + (size_t)xf_s * rn.stage_bytes, xf_tile, xf_tt, xf_nimg, xf_plane0, chunks, ptid, NPe);
                            }
                            if (tid == 0 && it < 6) trace_mark(p, 14 + 8 * (int)it);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            mbar_arrive(FULL(xf_s));
                            if (tid == 0 && it < 6) trace_mark(p, 13 + 8 * (int)it);
                        }
                        xf_valid = true; xf_s = s; xf_tile = tile; xf_tt = tt; xf_nimg = n_img; xf_plane0 = plane0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) xf_sc[j] = xf_sc_new[j];
                        xf_scn = xf_scn_new;
                    }
                } else {
                    if (LOADER == LD_STEM_U8) {
                        // uint8 NCHW source, C = 3, K = 32: one pair of horizontally adjacent output pixels per thread
                        // iteration.  Per (row, channel) two aligned 32-bit loads cover the five source bytes of the
                        // pair; bytes become EXACT bf16 integers 0..255 (PRMT into the 2^23 magic float, subtract, take
                        // the high half) and the 1/255 of BasePredictor.preprocess is applied to the fp32 accumulator in
                        // the epilogue (out_scale), which is closer to the fp32 reference than rounding u/255 to bf16.
                        const size_t plane_sz = (size_t)p.stem_H * p.stem_W;
                        const uint8_t* src8 = reinterpret_cast<const uint8_t*>(p.stem_src);
                        const uint32_t rows = 128u * rn.MB;
                        // the 18 source words of output-pixel pair pp of tile tl
                        auto stem_load = [&](uint32_t tl, uint32_t pp, uint32_t (&wa)[9], uint32_t (&wb)[9]) {
                            const uint32_t g = tl * rows + 2 * pp;
#pragma unroll
                            for (int j = 0; j < 9; ++j) { wa[j] = 0u; wb[j] = 0u; }
                            if (g < p.M_total) {
                                const uint32_t n = fdiv(g, p.d_HW);
                                const uint32_t rem = g - n * (uint32_t)(p.H * p.W);
                                const int ho = (int)fdiv(rem, p.d_Wo), wo = (int)rem - ho * p.W;   // wo is even
                                const uint8_t* base = src8 + (size_t)n * 3 * plane_sz + 2 * wo;
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    const int hi = 2 * ho + dy - 1;
                                    if (hi >= 0 && hi < p.stem_H) {
#pragma unroll
                                        for (int c = 0; c < 3; ++c) {
                                            const uint32_t* rp = reinterpret_cast<const uint32_t*>(base + c * plane_sz + (size_t)hi * p.stem_W);
                                            wb[dy * 3 + c] = __ldg(rp);
                                            if (wo > 0) wa[dy * 3 + c] = __ldg(rp - 1);
                                        }
                                    }
                                }
                            }
                        };
                        // When one pass of the producers covers the tile (rows / 2 <= NP), the words of the NEXT tile are
                        // requested as soon as this tile's have been converted, so that their latency (~1.5 us, the whole
                        // per-tile time of the unpipelined loader) overlaps the wait for the ring slot and the barrier.
                        const bool one_pass = rows / 2 <= (uint32_t)NP;
                        if (one_pass && it == 0 && (uint32_t)ptid < rows / 2) stem_load(tile, ptid, stem_wa, stem_wb);
                        for (uint32_t pp = ptid; pp < rows / 2; pp += NP) {
                            const uint32_t pos = 2 * pp;
                            uint32_t (&wa)[9] = stem_wa;
                            uint32_t (&wb)[9] = stem_wb;
                            if (!one_pass) stem_load(tile, pp, wa, wb);
#pragma unroll
                            for (int px = 0; px < 2; ++px) {
                                float f[32];
#pragma unroll
                                for (int j = 27; j < 32; ++j) f[j] = 0.f;
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                                    for (int c = 0; c < 3; ++c) {
                                        // low three bytes = source columns 2*wo'-1, 2*wo', 2*wo'+1 of output column wo' = wo + px
                                        const uint32_t x3 = px == 0 ? __funnelshift_r(wa[dy * 3 + c], wb[dy * 3 + c], 24) : (wb[dy * 3 + c] >> 8);
#pragma unroll
                                        for (int dx = 0; dx < 3; ++dx)
                                            f[(dy * 3 + dx) * 3 + c] = __uint_as_float(__byte_perm(x3, 0x4B000000u, 0x7650 | dx)) - 8388608.0f;
                                    }
                                }
#pragma unroll
                                for (int pll = 0; pll < 4; ++pll) {
                                    uint4 o;   // exact small integers: the bf16 value is the float's high half
                                    o.x = __byte_perm(__float_as_uint(f[pll * 8 + 0]), __float_as_uint(f[pll * 8 + 1]), 0x7632);
                                    o.y = __byte_perm(__float_as_uint(f[pll * 8 + 2]), __float_as_uint(f[pll * 8 + 3]), 0x7632);
                                    o.z = __byte_perm(__float_as_uint(f[pll * 8 + 4]), __float_as_uint(f[pll * 8 + 5]), 0x7632);
                                    o.w = __byte_perm(__float_as_uint(f[pll * 8 + 6]), __float_as_uint(f[pll * 8 + 7]), 0x7632);
                                    *reinterpret_cast<uint4*>(sA + ((uint32_t)pll * rn.pstride16 + pos + px) * 16u) = o;
                                }
                            }
                            if (one_pass && tile + gridDim.x < tiles) stem_load(tile + gridDim.x, pp, wa, wb);
                        }
                    } else if (LOADER == LD_DCN && pl.PS % (p.dcn_cin / 8) == 0) {
                        // DCNv2 sampling, one (position, tap) item per thread iteration: offset / mask / bilinear weights
                        // are decoded once and reused for all channel groups of the tap (mask folded into the weights,
                        // out-of-image corners get weight 0 and a clamped address, so the loop is branch-free)
                        const int cgs = p.dcn_cin / 8;
                        const int tap0 = plane0 / cgs, ntap = pl.PS / cgs;
                        const uint32_t rows = 128u * rn.MB;
                        const uint32_t items = rows * (uint32_t)ntap;
                        // The thread's items are a dependent chain of global-load latencies (offsets -> corners -> store);
                        // the offsets / mask of the NEXT item are requested while this one is sampled, and the corner loads
                        // of two channel groups are in flight together: 2 instead of 1 + cgs latencies per item.
                        auto dcn_raw = [&](uint32_t e, float& dy, float& dx, float& m) {
                            const uint32_t tl = e / rows, pos = e - tl * rows;
                            const int tap = tap0 + (int)tl;
                            const uint32_t g = tile * rows + pos;
                            dy = dx = m = 0.f;
                            if (e < items && g < p.M_total) {
                                const __nv_bfloat16* ofp = p.dcn_off + (size_t)g * p.off_cs + 2 * tap;
                                dy = __bfloat162float(ofp[0]); dx = __bfloat162float(ofp[1]);
                                m = __bfloat162float(p.dcn_mask[(size_t)g * p.mask_cs + tap]);
                            }
                        };
                        float ndy, ndx, nm;
                        dcn_raw(ptid, ndy, ndx, nm);
                        for (uint32_t e = ptid; e < items; e += NP) {
                            const uint32_t tl = e / rows, pos = e - tl * rows;   // consecutive threads -> consecutive pixels
                            const int tap = tap0 + (int)tl;
                            const uint32_t g = tile * rows + pos;
                            float wgt[4] = {0.f, 0.f, 0.f, 0.f};
                            const __nv_bfloat16* cp[4] = {p.x, p.x, p.x, p.x};
                            const float dy = ndy, dx = ndx;
                            float m = nm;
                            dcn_raw(e + NP, ndy, ndx, nm);
                            if (g < p.M_total) {
                                const uint32_t n = fdiv(g, p.d_HW);
                                const uint32_t rem = g - n * (uint32_t)(p.H * p.W);
                                const int hq = (int)fdiv(rem, p.d_W), wq = (int)rem - hq * p.W;
                                if (p.mask_logit) m = sigmoidf_(m);
                                const float py = (float)(hq + tap / 3 - 1) + dy, px = (float)(wq + tap % 3 - 1) + dx;
                                if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
                                    const int y0 = (int)floorf(py), x0 = (int)floorf(px);
                                    const float ly = py - (float)y0, lx = px - (float)x0;
                                    const float hy = 1.f - ly, hx = 1.f - lx;
                                    const bool y0v = y0 >= 0, y1v = y0 + 1 <= p.H - 1, x0v = x0 >= 0, x1v = x0 + 1 <= p.W - 1;
                                    wgt[0] = (y0v && x0v) ? hy * hx * m : 0.f;
                                    wgt[1] = (y0v && x1v) ? hy * lx * m : 0.f;
                                    wgt[2] = (y1v && x0v) ? ly * hx * m : 0.f;
                                    wgt[3] = (y1v && x1v) ? ly * lx * m : 0.f;
                                    const int yc0 = max(y0, 0), yc1 = min(y0 + 1, p.H - 1), xc0 = max(x0, 0), xc1 = min(x0 + 1, p.W - 1);
                                    const __nv_bfloat16* xn = p.x + (size_t)n * p.H * p.W * p.x_cs;
                                    cp[0] = xn + (size_t)(yc0 * p.W + xc0) * p.x_cs;
                                    cp[1] = xn + (size_t)(yc0 * p.W + xc1) * p.x_cs;
                                    cp[2] = xn + (size_t)(yc1 * p.W + xc0) * p.x_cs;
                                    cp[3] = xn + (size_t)(yc1 * p.W + xc1) * p.x_cs;
                                }
                            }
                            unsigned char* dst = sA + ((uint32_t)((tap - tap0) * cgs) * rn.pstride16 + pos) * 16u;
                            for (int cg0 = 0; cg0 < cgs; cg0 += 2) {
                                uint4 v[2][4];
                                if (p.dcn_wide && cg0 + 1 < cgs) {
                                    // LDG.E.256: both channel groups of a corner are one 32-byte sector -- two 16-byte loads
                                    // request that sector twice, and this gather is bound by L1 sector throughput (ncu: 29
                                    // sectors per request, l1tex 72 % of peak)
#pragma unroll
                                    for (int c4 = 0; c4 < 4; ++c4)
                                        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                                     : "=r"(v[0][c4].x), "=r"(v[0][c4].y), "=r"(v[0][c4].z), "=r"(v[0][c4].w),
                                                       "=r"(v[1][c4].x), "=r"(v[1][c4].y), "=r"(v[1][c4].z), "=r"(v[1][c4].w)
                                                     : "l"(cp[c4] + cg0 * 8));
                                } else {
#pragma unroll
                                    for (int u = 0; u < 2; ++u)
#pragma unroll
                                        for (int c4 = 0; c4 < 4; ++c4)
                                            v[u][c4] = __ldg(reinterpret_cast<const uint4*>(cp[c4] + min(cg0 + u, cgs - 1) * 8));
                                }
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    if (cg0 + u >= cgs) break;
                                    float f[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) f[j] = 0.f;
#pragma unroll
                                    for (int c4 = 0; c4 < 4; ++c4) {
                                        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v[u][c4]);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) {
                                            const float2 t = __bfloat1622float2(h[j]);
                                            f[2 * j] = fmaf(wgt[c4], t.x, f[2 * j]);
                                            f[2 * j + 1] = fmaf(wgt[c4], t.y, f[2 * j + 1]);
                                        }
                                    }
                                    uint4 o;
                                    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                                    for (int j = 0; j < 4; ++j) oh[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                                    *reinterpret_cast<uint4*>(dst + (size_t)(cg0 + u) * rn.pstride16 * 16u) = o;
                                }
                            }
                        }
                    } else {
                        // U2_MLP chunks in flight per thread: loads first, then stores
                        for (uint32_t e0 = ptid; e0 < chunks; e0 += NP * U2_MLP) {
                            uint4 v[U2_MLP];
                            uint32_t dsto[U2_MLP];
#pragma unroll
                            for (int u = 0; u < U2_MLP; ++u) {
                                const uint32_t e = e0 + u * NP;
                                dsto[u] = 0xffffffffu;
                                if (e < chunks) {
                                    const uint32_t rest = fdiv(e, p.d_ps);
                                    const int pll = (int)(e - rest * pl.PS);
                                    uint32_t pos = rest, par = 0;
                                    if (MODE == 2) { par = fdiv(rest, p.d_P); pos = rest - par * rn.P; }
                                    if (LOADER == LD_STEM_GEN) v[u] = stem_chunk(p, tile * (128u * rn.MB) + pos, plane0 + pll);
                                    else if (LOADER == LD_DCN) v[u] = dcn_chunk(p, tile * (128u * rn.MB) + pos, plane0 + pll);
                                    else v[u] = make_uint4(0u, 0u, 0u, 0u);
                                    dsto[u] = ((uint32_t)pll * rn.pstride16 + par * rn.P + pos) * 16u;
                                }
                            }
#pragma unroll
                            for (int u = 0; u < U2_MLP; ++u)
                                if (dsto[u] != 0xffffffffu) *reinterpret_cast<uint4*>(sA + dsto[u]) = v[u];
                        }
                    }
                    if (rn.w_slice_bytes) cp_async_wait_all();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(FULL(s));
                }
                if (tid == 0 && it < 6) trace_mark(p, 8 + 8 * (int)it);
                if (++s == rn.S) { s = 0; ph ^= 1; }
            }
        }
        if (LOADER == LD_XFORM && xf_valid) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            if (p_active) {
                                if (ps_divides) xform_stage_fixed<MODE>(p, sStage + (size_t)xf_s * rn.stage_bytes, xf_tile, xf_tt, xf_nimg,
                                                                        xf_plane0 + (int)pll_fix, pll_fix, pos_fix, pstep, xf_sc, xf_scn);
                                else xform_stage<MODE>(p, sStage + (size_t)xf_s * rn.stage_bytes, xf_tile, xf_tt, xf_nimg, xf_plane0, chunks, ptid, NPe);
                            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(FULL(xf_s));
        }
    } else if (warp < EPI0) {
        // =============================================================== MMA issuers (warps MMAW .. MMAW + NMW - 1)
        const int mw = warp - MMAW;   // this warp owns the chains (row block mb, partial accumulator h) with (mb * KS + h) % NMW == mw
        // The whole warp walks the pipeline in uniform control flow and computes the (warp-uniform) descriptors; only the
        // tcgen05 instructions themselves sit under the elect.sync predicate.  tools/ubench/umma_issue.cu: with the
        // descriptor arithmetic in per-thread registers (a table in shared memory, the loop inside `if (elect)`) every
        // K = 16 step costs ~150 cycles of LDS / R2UR latency before its first UTCHMMA (213 cycles per instruction at two
        // row blocks); on the uniform datapath a four-row-block step issues at ~58 cycles per instruction.
        // instruction descriptor: D = f32, A = bf16, B = bf16 or f16, both K-major, N = Nc, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | ((p.w_f16 ? 0u : 1u) << 10) | ((uint32_t)(pl.Nc >> 3) << 17) | ((128u >> 4) << 24);
        const bool lead = elect_one();
        if (rn.wres_bytes) mbar_wait(WREADY, 0);
        uint32_t it = 0, ti = 0;
        int s = 0;
        uint32_t ph = 0;
        const uint32_t stage0 = s_u32(sStage), wres0 = s_u32(sWres);
        const uint64_t bdesc_t = mk_desc(0u, (uint32_t)pl.Nc * 16u, 128u);   // B: chunk c at c * Nc * 16 bytes, LBO = Nc * 16
        const int kmask = rn.KS - 1;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
            const int a = rn.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = rn.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            mbar_wait(ACCEMPTY(a), aphase ^ 1);
            for (int ks = 0; ks < nks; ++ks, ++it) {
                mbar_wait(FULL(s), ph);
                // the stage was written through the generic proxy (cp.async / st.shared by the producers)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = stage0 + (uint32_t)s * rn.stage_bytes;
                const uint32_t w0 = rn.w_slice_bytes ? a0 + rn.a_bytes : wres0 + (uint32_t)ks * (uint32_t)(w_slice_elems * 2);
                const uint32_t d0 = tmem_base + (uint32_t)(a * rn.KS * rn.MB * pl.Nc);
                const uint64_t abase = (uint64_t)(a0 >> 4);
                const uint64_t bdesc0 = bdesc_t + (uint64_t)(w0 >> 4);
                // common case (one issuing warp, no K-split): one copy of the K loop per tile height, the MB instructions of
                // a step under ONE branch on the elected lane
                auto kloop = [&](auto mbc) {
                    constexpr int MBK = decltype(mbc)::value;
                    uint64_t bdesc = bdesc0;
                    const uint32_t acc0 = ks ? 1u : 0u;
#pragma unroll 2
                    for (int i = 0; i < pl.nmma_s; ++i, bdesc += (uint64_t)(2 * pl.Nc)) {
                        const uint64_t ad0 = p.adesc[i] + abase;
                        const uint32_t acc = i ? 1u : acc0;            // the first step of the tile overwrites
                        if (lead) {
#pragma unroll
                            for (int mb = 0; mb < MBK; ++mb)   // 2048 B of A and Nc accumulator columns per row block
                                asm volatile(
                                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                    ::"r"(d0 + (uint32_t)(mb * pl.Nc)), "l"(ad0 + (uint64_t)(mb * 128)), "l"(bdesc), "r"(idesc), "r"(acc));
                        }
                    }
                };
                if (NMW == 1 && rn.KS == 1 && rn.MB == 4) kloop(std::integral_constant<int, 4>());
                else if (NMW == 1 && rn.KS == 1 && rn.MB == 2) kloop(std::integral_constant<int, 2>());
                else if (NMW == 1 && rn.KS == 1 && rn.MB == 1) kloop(std::integral_constant<int, 1>());
                else {
                    uint64_t bdesc = bdesc0;
                    for (int i = 0; i < pl.nmma_s; ++i, bdesc += (uint64_t)(2 * pl.Nc)) {
                        const uint64_t ad0 = p.adesc[i] + abase;
                        const int g = ks * pl.nmma_s + i;                  // K = 16 step of this tile
                        const uint32_t acc = g >= rn.KS ? 1u : 0u;         // the first step of every partial accumulator overwrites
                        const uint32_t dh = d0 + (uint32_t)((g & kmask) * rn.MB * pl.Nc);
                        for (int mb = 0; mb < rn.MB; ++mb) {
                            if (NMW > 1 && ((mb * rn.KS + (g & kmask)) & (NMW - 1)) != mw) continue;
                            if (lead)
                                asm volatile(
                                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                    ::"r"(dh + (uint32_t)(mb * pl.Nc)), "l"(ad0 + (uint64_t)(mb * 128)), "l"(bdesc), "r"(idesc), "r"(acc));
                        }
                    }
                }
                if (lead) {
                    umma_commit(EMPTY(s));                       // smem stage may be refilled once these MMAs retire
                    if (ks == nks - 1) umma_commit(ACCFULL(a));  // accumulators complete
                    if (mw == 0 && it < 6) trace_mark(p, 9 + 8 * (int)it);
                }
                __syncwarp();
                if (++s == rn.S) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================================================== epilogue (warps EPI0 .. 19)
        // TMEM lane quadrant = warp % 4; the NEW/4 warps of a quadrant take the (row block, 32-column) units round
        // robin.  A unit is read from TMEM one output row per lane, converted, transposed through a swizzled 2 KB
        // shared-memory tile and written out with 8 rows x 64 contiguous bytes per store instruction.
        pdl_wait();
        const int ew = warp - EPI0;
        const int quad = warp & 3, sub = ew >> 2;
        constexpr int NSUB = NEW / 4;
        const int ncch = (pl.Nc + 31) / 32;
        // 2 KB staging tile per warp, 512-byte aligned: chunk c of row r at slot c ^ ((r >> 1) & 3) is then exactly the
        // SWIZZLE_64B pattern of a TMA tensor map (address bits [4:5] ^= bits [7:8]), so the tile can be stored by TMA
        const uint32_t stg32 = ((s_u32(sOut) + 511u) & ~511u) + (uint32_t)ew * 2048u;
        const int srow = lane >> 2, schunk = lane & 3;    // store phase: row within a group of 8, 16-byte chunk
        const uint32_t st_wr = stg32 + lane * 64, sw_wr = (uint32_t)((lane >> 1) & 3);
        uint32_t st_rd[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int row = g * 8 + srow;
            st_rd[g] = stg32 + row * 64 + ((schunk ^ ((row >> 1) & 3)) << 4);
        }
        const bool trw = p.trace && ew == 0 && lane == 0;
        uint32_t ti = 0;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
            const int a = rn.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = rn.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            mbar_wait(ACCFULL(a), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trw && ti < 6) trace_mark(p, 10 + 8 * (int)ti);
            // 16-column tiles (Cout <= 16: the stem, the narrow 3x3 layers): the accumulators of row blocks mb and mb + 1
            // are adjacent 16-column groups in TMEM, so ONE 32-column unit covers both -- columns 0..15 are channels
            // 0..15 of row block mb, columns 16..31 the same channels of row block mb + 1.  The per-unit cost of this
            // latency-bound loop is the same for 16 and 32 columns, so pairing halves the epilogue time of these layers.
            const bool pair = !STATS && p.pair_ok && pl.Nc == 16 && (rn.MB & 1) == 0 && (p.residual == nullptr || p.Cout <= 8) && p.row_scale == nullptr;   // (measured: a loss for 16-channel residual layers)
            const int mbstep = pair ? 2 : 1;
            int u = sub;                                           // units are numbered mg * ncch + cc
            for (int mg = 0; mg * mbstep < rn.MB; ++mg) {
                const int mb = mg * mbstep;
                if (u >= (mg + 1) * ncch) continue;                // no unit of this row block (pair) is ours
                const int opix = out_pixel2(p, tile, (uint32_t)(mb * 128 + quad * 32 + lane));
                const int opixB = pair ? out_pixel2(p, tile, (uint32_t)((mb + 1) * 128 + quad * 32 + lane)) : -1;
                const bool any_row = __any_sync(0xffffffffu, opix >= 0 || opixB >= 0);
                uint32_t skey = 0xffffffffu;   // fused statistics: (image << 4) | adaptive-pool window mask of this lane's row
                if (STATS && opix >= 0) {
                    const uint32_t n = fdiv((uint32_t)opix, p.d_oHW);
                    uint32_t mask = 0;
                    if (p.st_Q == 5) {
                        const uint32_t rem = (uint32_t)opix - n * (uint32_t)(p.Ho * p.Wo);
                        const int h = (int)fdiv(rem, p.d_oW), w = (int)rem - h * p.Wo;
                        const uint32_t top = h < p.st_h0e, bot = h >= p.st_h1b, lef = w < p.st_w0e, rig = w >= p.st_w1b;
                        mask = (top & lef) | ((top & rig) << 1) | ((bot & lef) << 2) | ((bot & rig) << 3);
                    }
                    skey = (n << 4) | mask;
                }
                __nv_bfloat16* yrow[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int orowA = __shfl_sync(0xffffffffu, opix, g * 8 + srow);
                    const int orowB = __shfl_sync(0xffffffffu, opixB, g * 8 + srow);
                    const int orow = (pair && schunk >= 2) ? orowB : orowA;   // paired unit: chunks 2, 3 are the second row block
                    yrow[g] = orow >= 0 ? p.y + (size_t)orow * p.y_cs : nullptr;
                }
                for (; u < (mg + 1) * ncch; u += NSUB) {
                    const int cl = (u - mg * ncch) * 32;          // first column of the unit within this CTA's Nc
                    const int co0 = ns * pl.Nc + cl;
                    if (!any_row || co0 >= p.Cout) continue;
                    const int nv = pair ? 32 : min(32, pl.Nc - cl);   // 32, or 16 for the last unit when Nc % 32 == 16
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                           (uint32_t)(a * rn.KS * rn.MB * pl.Nc + mb * pl.Nc + cl);
                    uint32_t r[32], pk[16];
                    const bool tr = trw && ti == 1;
                    long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
                    if (tr) tc0 = clock64();
                    if (nv == 32) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        for (int h = 1; h < rn.KS; ++h) {
                            tmem_ld_add16(taddr + (uint32_t)(h * rn.MB * pl.Nc), r);
                            tmem_ld_add16(taddr + (uint32_t)(h * rn.MB * pl.Nc + 16), r + 16);
                        }
                        if (tr) tc1 = clock64();
                        epi_math<32>(p, r, sBias, cl, co0, opix, pk, pair, opixB);
                    } else {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        for (int h = 1; h < rn.KS; ++h) tmem_ld_add16(taddr + (uint32_t)(h * rn.MB * pl.Nc), r);
                        if (tr) tc1 = clock64();
                        epi_math<16>(p, r, sBias, cl, co0, opix, pk);
#pragma unroll
                        for (int j = 8; j < 16; ++j) pk[j] = 0u;
                    }
                    const bool tma_on = MODE == 0 && p.tma_store && !pair;   // uniform per launch
                    // a 16-column unit (the tail of a column split) keeps the LSU stores: the 32-channel box would spill
                    // into the next split's channels (clipping only happens at Cout)
                    const bool tma = tma_on && nv == 32;
                    if (tma_on) {   // the previous unit's TMA store has read the staging tile
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        __syncwarp();
                    }
                    // row `lane` -> staging: 64 bytes per row, 16-byte chunk c at slot c ^ ((row >> 1) & 3)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_wr + ((c ^ sw_wr) << 4)), "r"(pk[4 * c]),
                                     "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                    if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (tr) tc2 = clock64();
                    if (tma) {
                        // one TMA tensor store per unit: the engine reads the swizzled tile and writes 32 rows x 64 bytes
                        // (clipped at Cout and at the end of the map / image), off the LSU path the producers' LDGSTS use
                        if (lane == 0) {
                            const uint32_t r0 = (uint32_t)(mb * 128 + quad * 32);
                            if (rn.per_img) {
                                const uint32_t n = fdiv(tile, p.d_tpi);
                                const uint32_t q = (tile - n * rn.tiles_per_img) * (128u * rn.MB) + r0;
                                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                             ::"l"(&p.ymap), "r"(co0), "r"(q), "r"(n), "r"(stg32) : "memory");
                            } else {
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                             ::"l"(&p.ymap), "r"(co0), "r"(tile * (128u * rn.MB) + r0), "r"(stg32) : "memory");
                            }
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else {
                    const int c8 = co0 + (pair ? (schunk & 1) : schunk) * 8;   // first output channel of this lane's chunk
                    const bool chunk_on = schunk * 8 < nv && c8 < p.Cout;
                    const bool full8 = c8 + 8 <= p.Cout && p.y_vec;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 o;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(st_rd[g]));
                        if (yrow[g] != nullptr && chunk_on) {
                            __nv_bfloat16* yp = yrow[g] + c8;
                            if (full8) *reinterpret_cast<uint4*>(yp) = o;
                            else {
                                const __nv_bfloat16* oh = reinterpret_cast<const __nv_bfloat16*>(&o);
                                for (int j = 0; j < 8 && c8 + j < p.Cout; ++j) yp[j] = oh[j];
                            }
                        }
                    }
                    }
                    if (STATS)
                        epi_stats(p.st_acc + (size_t)(tile % (uint32_t)p.st_R) * p.st_rs, p.st_Q, p.st_sq, p.st_tot, p.Cout, stg32, skey,
                                  lane, nv, co0, p.Cout);
                    __syncwarp();
                    if (tr) {
                        tc3 = clock64();
                        unsigned long long* t = p.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + 56;
                        t[0] += 1; t[1] += (unsigned long long)(tc1 - tc0); t[2] += (unsigned long long)(tc2 - tc1);
                        t[3] += (unsigned long long)(tc3 - tc2);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            // the accumulator buffer is free once this warp's tcgen05.ld's have completed (tcgen05.wait::ld above + the
            // fence); a relaxed arrival does not wait for the tile's global stores / statistics reductions to drain
            if (lane == 0) mbar_arrive_relaxed(ACCEMPTY(a));
            if (trw && ti < 6) trace_mark(p, 11 + 8 * (int)ti);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last TMA stores are complete before the CTA exits
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) trace_mark(p, 3);
    if (warp == MMAW) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)rn.tmem_cols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------- host side
static bool plan2_for(int Cin, int Cout, int k, int stride, int N, int H, int W, Plan2& pl, Run2& rn, int& Ho, int& Wo,
                      bool per_img_w = false) {
    pl = make_plan2(Cin, Cout, k, stride);
    if (!pl.ok) return false;
    const int pad = k / 2;
    Ho = (H + 2 * pad - k) / stride + 1;
    Wo = (W + 2 * pad - k) / stride + 1;
    if ((long long)N * H * W >= (1LL << 30) || (long long)N * Ho * Wo >= (1LL << 30)) return false;
    if (!make_run2(pl, N, H, W, Ho, Wo, rn, per_img_w)) return false;
    return rn.tiles < (1LL << 30);
}

static void fill_divs(P2& p) {
    p.d_ps = make_fastdiv((uint32_t)p.pl.PS);
    p.d_P = make_fastdiv((uint32_t)p.rn.P);
    p.d_Wq = make_fastdiv((uint32_t)p.rn.Wq);
    p.d_HW = make_fastdiv((uint32_t)(p.H * p.W));
    p.d_tpi = make_fastdiv((uint32_t)p.rn.tiles_per_img);
    p.d_W = make_fastdiv((uint32_t)p.W);
    p.d_cgs = make_fastdiv((uint32_t)(p.dcn_cin > 0 ? p.dcn_cin / 8 : 1));
    p.d_Wo = make_fastdiv((uint32_t)p.W);
    p.d_oHW = make_fastdiv((uint32_t)(p.Ho * p.Wo));
    p.d_oW = make_fastdiv((uint32_t)p.Wo);
}

// A descriptors of the K = 16 steps of a (slice, 128-row block): chunk c = (tap t, plane-in-slice pll); an instruction
// takes chunks 2i and 2i + 1 (LBO = their distance: the next plane, or -- Cin = 8 -- the next tap)
static void fill_adesc(P2& p) {
    const Plan2& pl = p.pl;
    const Run2& rn = p.rn;
    auto off = [&](int c) -> uint32_t {
        const int t = c / pl.PS, pll = c - t * pl.PS;
        int shift;
        if (pl.mode == 0) shift = 0;
        else if (pl.mode == 1) shift = rn.halo + (pl.tap_dy[t] - 1) * rn.Wq + (pl.tap_dx[t] - 1);
        else shift = (pl.tap_dy[t] >> 1) * rn.Wq + (pl.tap_dx[t] >> 1);
        return ((uint32_t)pll * rn.pstride16 + (uint32_t)pl.tap_par[t] * rn.P + shift) * 16u;
    };
    for (int i = 0; i < pl.nmma_s && i < U2_MAX_MMA; ++i) {
        const int c0 = 2 * i, c1 = 2 * i + 1;
        const uint32_t o0 = off(c0);
        const uint32_t lbo = (c1 < pl.taps * pl.PS) ? (off(c1) - o0) : 16u;  // dummy chunk: its weights are zero
        p.adesc[i] = (uint64_t)((o0 >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
    }
}

static unsigned long long* g_trace = nullptr;
// Library options, set through the C ABI (mgdt_set_option; the Python layer forwards MGDT_* environment variables once
// at load time): no getenv and no latched statics in the launch path.
static int g_force_split = -1;     // "conv_split" 0/1/2: override the producer/epilogue warp split (-1 = cost model)
static int g_tma_store = 1;        // "conv_tma_store": TMA tensor stores of 1x1 epilogue units
static int g_pair = 1;             // "conv_pair": paired 16-column epilogue units
static int g_use_tma_loads = 1;    // "conv_tma_load": TMA-fed kernel for transform-free 1x1 layers (conv_tma1x1.cuh)
static int g_tma_stats = 0;        // "conv_tma_stats": also for layers with fused output statistics (measured slower than the 16-epilogue-warp split of conv_umma2_kernel: 96->384 GELU + sum of squares 41 -> 55 us)

#include "conv_tma1x1.cuh"
#include "conv_tma3x3.cuh"

template <int MODE, int LOADER, int SPLIT, int STATS>
static int launch2k(const P2& p, dim3 grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma2_kernel<MODE, LOADER, SPLIT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_MAX_SMEM);
    if (e != cudaSuccess) return set_error(-EIO, "conv_umma2: smem attr: %s", cudaGetErrorString(e));
    launch_k(conv_umma2_kernel<MODE, LOADER, SPLIT, STATS>, dim3(grid), dim3(U2_THREADS), p.rn.smem_total, s, p);
    MGDT_LAUNCH_CHECK("conv_umma2");
    return 0;
}

// The statistics epilogue (STATS = 1) is instantiated only for the transform-free 1x1 / 3x3 stride-1 loaders -- the
// producers of every map the model takes statistics of (MSPA convs[-1], ConvNeXt pwconv1, Conv_GN) -- so that all other
// instantiations keep their code unchanged (+ the DCNv2 loader, whose output feeds a GroupNorm).
template <int MODE, int LOADER> constexpr bool stats_variant() { return (LOADER == LD_ASYNC && (MODE == 0 || MODE == 1)) || LOADER == LD_DCN; }

template <int MODE, int LOADER, int SPLIT>
static int launch2t(const P2& p, dim3 grid, cudaStream_t s) {
    if (p.st_acc) {
        if constexpr (stats_variant<MODE, LOADER>()) return launch2k<MODE, LOADER, SPLIT, 1>(p, grid, s);
        else return set_error(-ENOTSUP, "conv_umma2: fused statistics are not built for this loader");
    }
    return launch2k<MODE, LOADER, SPLIT, 0>(p, grid, s);
}

template <int MODE, int LOADER>
static int launch2s(const P2& p, int split, dim3 grid, cudaStream_t s) {
    if (split == 0) return launch2t<MODE, LOADER, 0>(p, grid, s);
    if (split == 1) return launch2t<MODE, LOADER, 1>(p, grid, s);
    if (split == 3 && LOADER == LD_ASYNC) return launch2t<MODE, LD_ASYNC, 3>(p, grid, s);
    return launch2t<MODE, LOADER, 2>(p, grid, s);
}

// Producer / epilogue warp split of the cp.async loaders.  Measured on the full model (profiles/, per-shape sweep of
// MGDT_CONV_SPLIT): a warp sustains only one 512-byte LDGSTS per ~200 cycles here, so 11 producer warps win almost
// everywhere; only tiles with many 32-column epilogue units (wide Cout) want the epilogue-heavy splits.
static int pick_split(const P2& p, bool xform) {
    (void)xform;
    const int units = p.rn.MB * ((p.pl.Nc + 31) / 32);
    // the statistics epilogue is ~40 % longer per unit: 1x1 layers that carry it want the epilogue-heavy split already
    // at four units (32->32 @160^2: 62 -> 49 us, 80->64 @80^2: 41 -> 34 us with 3 / 16 warps)
    if (p.st_acc && p.pl.mode == 0) return units >= 6 ? 1 : 0;
    // wide-K layers (512->256) keep more producers even with eight units (23 -> 18 us)
    if (units >= 8) return p.Cin >= 256 ? 1 : 0;
    return units >= 6 ? 1 : 2;
}

static int launch2(P2& p, cudaStream_t s) {
    fill_divs(p);
    fill_adesc(p);
    p.trace = g_trace;
    // TMA store map of the output (mode 0 layers with 16-byte aligned rows; the paired 16-column units keep the LSU path)
    p.tma_store = 0;
    if (g_tma_store && p.pl.mode == 0 && p.y_vec && p.Cout >= 8) {
        const unsigned long long HW = (unsigned long long)p.Ho * p.Wo;
        cuuint64_t dims[3] = {(cuuint64_t)p.Cout, p.rn.per_img ? HW : (cuuint64_t)p.M_total, (cuuint64_t)p.N};
        cuuint64_t strides[2] = {(cuuint64_t)p.y_cs * 2, HW * (cuuint64_t)p.y_cs * 2};
        cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
        EncodeTiledFn encode = tensor_map_encoder();   // absent driver entry point: the LSU store path is used
        if (encode) {
            const CUresult r = encode(&p.ymap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p.rn.per_img ? 3 : 2, (void*)p.y, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            p.tma_store = r == CUDA_SUCCESS ? 1 : 0;
        }
    }
    p.pair_ok = g_pair;
    if (p.pl.mode == 0) {   // transform-free 1x1 layers: the TMA-fed kernel
        const int rc = try_launch_t1(p, s);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (p.pl.mode == 1 || p.pl.mode == 2) {   // transform-free 3x3 layers whose whole K extent fits one stage: the TMA-fed kernel
        const int rc = try_launch_t3(p, s);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (p.act_cols) return set_error(-ENOTSUP, "conv2d: act_cols needs the TMA-fed 1x1 kernel (see mgdt_conv2d_path)");
    const long long tiles = p.rn.tiles;
    int ctas = (int)(tiles < 148 ? tiles : 148);
    if (p.pl.nsplit > 1) ctas = (int)std::max(1LL, std::min(tiles, (long long)(148 / p.pl.nsplit)));
    const dim3 grid((unsigned)ctas, (unsigned)p.pl.nsplit);
    if (p.stem_src) {
        if (p.out_scale != 1.0f) return launch2t<0, LD_STEM_U8, 0>(p, grid, s);
        return launch2t<0, LD_STEM_GEN, 0>(p, grid, s);
    }
    if (p.dcn_off) return launch2t<0, LD_DCN, 0>(p, grid, s);
    const bool xform = p.pre_add || p.in_scale || p.pix_scale || p.in_relu;
    const int split = g_force_split >= 0 ? g_force_split : pick_split(p, xform);
    switch (p.pl.mode * 2 + (xform ? 1 : 0)) {
        case 0: return launch2s<0, LD_ASYNC>(p, split, grid, s);
        case 1: return launch2s<0, LD_XFORM>(p, split, grid, s);
        case 2: return launch2s<1, LD_ASYNC>(p, split, grid, s);
        case 3: return launch2s<1, LD_XFORM>(p, split, grid, s);
        case 4: return launch2s<2, LD_ASYNC>(p, split, grid, s);
        default: return launch2s<2, LD_XFORM>(p, split, grid, s);
    }
}

bool conv2d_umma_supported(const mgdt_conv_args* a) {
    if (!a->w_umma || a->dtype != MGDT_BF16 || a->kh != a->kw || a->pad != a->kh / 2) return false;
    // fused statistics: transform-free stride-1 loaders only (see stats_variant)
    if (a->stat_acc && (a->stride != 1 || a->pre_add || a->in_scale || a->pix_scale || a->in_relu)) return false;
    // per-image weights (w_umma = N packed images): transform-free loader only, nothing else scaled per (n, c)
    if (a->w_per_image && (a->pre_add || a->in_scale || a->pix_scale || a->in_relu)) return false;
    if (a->act_cols && (a->act_cols % 32 != 0 || a->kh != 1 || a->stride != 1 || a->stat_acc)) return false;
    Plan2 pl; Run2 rn; int Ho, Wo;
    if (!plan2_for(a->Cin, a->Cout, a->kh, a->stride, a->N, a->H, a->W, pl, rn, Ho, Wo, a->w_per_image != 0)) return false;
    if (((uintptr_t)a->x & 15) || (a->x_cs & 7)) return false;
    if (a->pre_add && (((uintptr_t)a->pre_add & 15) || (a->add_cs & 7))) return false;
    if (a->in_scale && ((uintptr_t)a->in_scale & 15)) return false;
    if ((uintptr_t)a->w_umma & 15) return false;
    return true;
}

static int fill_p2(const mgdt_conv_args* a, P2& p) {
    int Ho, Wo;
    if (!plan2_for(a->Cin, a->Cout, a->kh, a->stride, a->N, a->H, a->W, p.pl, p.rn, Ho, Wo, a->w_per_image != 0))
        return set_error(-EINVAL, "conv2d_umma: unsupported shape");
    p.w_img_elems = a->w_per_image ? (size_t)p.pl.nsplit * p.pl.nks * p.pl.nmma_s * 2 * p.pl.Nc * 8 : 0;
    p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w_umma;
    p.pre_add = (const __nv_bfloat16*)a->pre_add; p.pix_scale = (const __nv_bfloat16*)a->pix_scale;
    p.residual = (const __nv_bfloat16*)a->residual; p.bias = a->bias; p.in_scale = a->in_scale;
    p.y = (__nv_bfloat16*)a->y;
    p.row_scale = nullptr;
    if (p.pl.mode == 0 && p.pix_scale && !p.pre_add && !p.in_scale && !a->in_relu) {   // commutes with a 1x1 conv
        p.row_scale = p.pix_scale;
        p.pix_scale = nullptr;
    }
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.Ho = Ho; p.Wo = Wo;
    p.x_cs = a->x_cs; p.y_cs = a->y_cs; p.add_cs = a->add_cs; p.ps_cs = a->ps_cs; p.res_cs = a->res_cs;
    p.act = a->act; p.in_relu = a->in_relu; p.act_cols = a->act_cols;
    p.w_f16 = a->w_umma_f16; p.out_scale = 1.0f;
    p.y_vec = (((uintptr_t)a->y & 15) == 0 && (a->y_cs & 7) == 0) ? 1 : 0;
    p.res_vec = (a->residual && ((uintptr_t)a->residual & 15) == 0 && (a->res_cs & 7) == 0) ? 1 : 0;
    p.M_total = (unsigned)((long long)a->N * a->H * a->W);
    p.dcn_off = nullptr; p.dcn_mask = nullptr; p.off_cs = p.mask_cs = p.mask_logit = p.dcn_cin = 0; p.dcn_wide = 0;
    p.stem_src = nullptr; p.stem_u8 = p.stem_C = p.stem_H = p.stem_W = 0;
    p.st_acc = (double*)a->stat_acc; p.st_Q = a->stat_q; p.st_sq = a->stat_sq ? 1 : 0;
    p.st_h0e = (Ho + 1) / 2; p.st_h1b = Ho / 2; p.st_w0e = (Wo + 1) / 2; p.st_w1b = Wo / 2;
    p.st_R = a->stat_copies > 0 ? a->stat_copies : 1;
    p.st_rs = (long long)a->N * (p.st_Q + p.st_sq) * a->Cout;
    p.st_tot = (p.st_Q == 1 || (p.st_Q == 5 && ((Ho | Wo) & 1))) ? 1 : 0;
    return 0;
}

int conv2d_umma(const mgdt_conv_args* a, cudaStream_t s) {
    P2 p;
    const int rc = fill_p2(a, p);
    return rc < 0 ? rc : launch2(p, s);
}

// 2 = conv_umma2_kernel (cp.async-fed), 4 = conv1x1_tma_kernel (TMA-fed), for a layer conv2d_umma_supported() accepts
int conv2d_umma_path(const mgdt_conv_args* a) {
    P2 p;
    if (fill_p2(a, p) < 0) return 2;
    if (p.pl.mode == 0 && t1_eligible(p) && plan_t1(p, p.t1) && tensor_map_encoder()) return 4;
    if ((p.pl.mode == 1 || p.pl.mode == 2) && t3_eligible(p) && plan_t3(p, p.t3) && tensor_map_encoder()) return 5;
    return 2;
}

bool dcn_umma_supported(const void* x, int x_cs, const void* w_umma, int N, int H, int W, int Cin, int Cout) {
    if (!w_umma || Cin % 8 != 0 || ((uintptr_t)x & 15) || (x_cs & 7) || ((uintptr_t)w_umma & 15)) return false;
    Plan2 pl; Run2 rn; int Ho, Wo;
    return plan2_for(9 * Cin, Cout, 1, 1, N, H, W, pl, rn, Ho, Wo);
}

int dcn_umma(const void* x, int x_cs, const void* offset, int off_cs, const void* mask, int mask_cs, int mask_is_logit,
             const void* w_umma, int w_f16, void* y, int y_cs, int N, int H, int W, int Cin, int Cout, void* stat_acc, int stat_q,
             int stat_sq, int stat_copies, cudaStream_t s) {
    P2 p;
    int Ho, Wo;
    if (!plan2_for(9 * Cin, Cout, 1, 1, N, H, W, p.pl, p.rn, Ho, Wo)) return set_error(-EINVAL, "dcn_umma: unsupported shape");
    p.x = (const __nv_bfloat16*)x; p.w = (const __nv_bfloat16*)w_umma; p.pre_add = nullptr; p.pix_scale = nullptr;
    p.residual = nullptr; p.bias = nullptr; p.in_scale = nullptr; p.y = (__nv_bfloat16*)y; p.row_scale = nullptr;
    p.N = N; p.H = H; p.W = W; p.Cin = 9 * Cin; p.Cout = Cout; p.Ho = H; p.Wo = W;
    p.x_cs = x_cs; p.y_cs = y_cs; p.add_cs = p.ps_cs = p.res_cs = 0; p.act = MGDT_ACT_NONE; p.in_relu = 0; p.act_cols = 0;
    p.w_f16 = w_f16; p.out_scale = 1.0f;
    p.y_vec = (((uintptr_t)y & 15) == 0 && (y_cs & 7) == 0) ? 1 : 0;
    p.res_vec = 0;
    p.M_total = (unsigned)((long long)N * H * W);
    p.w_img_elems = 0;
    p.dcn_off = (const __nv_bfloat16*)offset; p.dcn_mask = (const __nv_bfloat16*)mask;
    p.off_cs = off_cs; p.mask_cs = mask_cs; p.mask_logit = mask_is_logit; p.dcn_cin = Cin;
    p.dcn_wide = ((((uintptr_t)x) & 31) == 0 && (x_cs & 15) == 0 && (Cin & 15) == 0) ? 1 : 0;
    p.stem_src = nullptr; p.stem_u8 = p.stem_C = p.stem_H = p.stem_W = 0;
    p.st_acc = (double*)stat_acc; p.st_Q = stat_acc ? stat_q : 0; p.st_sq = (stat_acc && stat_sq) ? 1 : 0;
    p.st_h0e = (H + 1) / 2; p.st_h1b = H / 2; p.st_w0e = (W + 1) / 2; p.st_w1b = W / 2;
    p.st_R = stat_copies > 0 ? stat_copies : 1;
    p.st_rs = (long long)N * (p.st_Q + p.st_sq) * Cout;
    p.st_tot = (p.st_Q == 1 || (p.st_Q == 5 && ((H | W) & 1))) ? 1 : 0;
    return launch2(p, s);
}

// Fused preprocess + stem conv (3x3, stride 2, pad 1, C <= 3..8 input channels): a 1x1 GEMM over the virtual
// K = round_up(9*C, 16) im2col built from the NCHW source while staging; output geometry (Ho, Wo) plays the
// role of the 1x1 conv's (H, W).
int stem_umma(const void* src, int src_is_u8, const void* w_umma, int w_f16, const float* bias, void* y, int y_cs, int N,
              int C, int H, int W, int Cout, int act, cudaStream_t s) {
    const int Kp = (9 * C + 15) / 16 * 16;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    P2 p;
    int ho2, wo2;
    if (!plan2_for(Kp, Cout, 1, 1, N, Ho, Wo, p.pl, p.rn, ho2, wo2)) return set_error(-EINVAL, "stem_umma: unsupported shape");
    p.x = nullptr; p.w = (const __nv_bfloat16*)w_umma; p.pre_add = nullptr; p.pix_scale = nullptr; p.residual = nullptr;
    p.bias = bias; p.in_scale = nullptr; p.y = (__nv_bfloat16*)y; p.row_scale = nullptr;
    p.N = N; p.H = Ho; p.W = Wo; p.Cin = Kp; p.Cout = Cout; p.Ho = Ho; p.Wo = Wo;
    p.x_cs = 0; p.y_cs = y_cs; p.add_cs = p.ps_cs = p.res_cs = 0; p.act = act; p.in_relu = 0; p.act_cols = 0;
    p.w_f16 = w_f16;
    // the fast uint8 loader stages exact integers and leaves the /255 to the epilogue
    const bool u8_fast = src_is_u8 && C == 3 && p.pl.PS == 4 && W % 4 == 0 && ((uintptr_t)src & 3) == 0;
    p.out_scale = u8_fast ? 1.0f / 255.0f : 1.0f;
    p.y_vec = (((uintptr_t)y & 15) == 0 && (y_cs & 7) == 0) ? 1 : 0;
    p.res_vec = 0;
    p.M_total = (unsigned)((long long)N * Ho * Wo);
    p.dcn_off = nullptr; p.dcn_mask = nullptr; p.off_cs = p.mask_cs = p.mask_logit = p.dcn_cin = 0; p.dcn_wide = 0;
    p.w_img_elems = 0;
    p.stem_src = src; p.stem_u8 = src_is_u8; p.stem_C = C; p.stem_H = H; p.stem_W = W;
    p.st_acc = nullptr; p.st_Q = p.st_sq = p.st_h0e = p.st_h1b = p.st_w0e = p.st_w1b = p.st_tot = 0; p.st_R = 1; p.st_rs = 0;
    return launch2(p, s);
}

}  // namespace mgdt

using namespace mgdt;

extern "C" void mgdt_debug_set_trace(void* buf) { mgdt::g_trace = (unsigned long long*)buf; }

namespace mgdt {
int conv_set_option(const char* name, int value) {
    if (!strcmp(name, "conv_mb")) g_force_mb = (value == 1 || value == 2 || value == 4) ? value : 0;
    else if (!strcmp(name, "conv_split")) g_force_split = (value >= 0 && value <= 3) ? value : -1;
    else if (!strcmp(name, "conv_tma_store")) g_tma_store = value ? 1 : 0;
    else if (!strcmp(name, "conv_pair")) g_pair = value ? 1 : 0;
    else if (!strcmp(name, "conv_tma_load")) g_use_tma_loads = value ? 1 : 0;
    else if (!strcmp(name, "conv_tma3x3")) g_use_tma3 = value ? 1 : 0;
    else if (!strcmp(name, "conv_tma3x3_s2")) g_use_tma3_s2 = value ? 1 : 0;
    else if (!strcmp(name, "conv_tma_stats")) g_tma_stats = value ? 1 : 0;
    else if (!strcmp(name, "conv_ksplit")) g_ksplit = value ? 1 : 0;
    else return 0;
    return 1;
}
}  // namespace mgdt

extern "C" int mgdt_stem_conv(const void* src, int src_is_u8, const void* w_umma, int w_umma_f16, const float* bias,
                              void* y, int y_cs, int N, int C, int H, int W, int Cout, int act, int dtype, void* stream) {
    MGDT_CHECK(src && w_umma && y, "stem_conv: null pointer");
    MGDT_CHECK(dtype == MGDT_BF16, "stem_conv: the fused tensor-core stem is bf16 only (use preprocess + conv2d for fp32)");
    MGDT_CHECK(N > 0 && C > 0 && C <= 8 && H > 1 && W > 1 && Cout > 0 && y_cs >= Cout, "stem_conv: bad shape");
    MGDT_CHECK(((uintptr_t)w_umma & 15) == 0, "stem_conv: packed weights must be 16-byte aligned");
    return stem_umma(src, src_is_u8, w_umma, w_umma_f16, bias, y, y_cs, N, C, H, W, Cout, act, (cudaStream_t)stream);
}

extern "C" size_t mgdt_conv_umma_packed_bytes(int Cin, int Cout, int k, int stride) {
    const Plan2 pl = make_plan2(Cin, Cout, k, stride);
    if (!pl.ok) return 0;
    return (size_t)pl.nsplit * pl.nks * pl.nmma_s * 2 * pl.Nc * 16;
}

extern "C" int mgdt_conv_umma_pack(const void* w_ohwi, int w_dtype, int Cin, int Cout, int k, int stride, int out_f16,
                                   void* packed, void* stream) {
    MGDT_CHECK(w_ohwi && packed, "conv_umma_pack: null pointer");
    MGDT_CHECK(w_dtype == MGDT_F32 || w_dtype == MGDT_BF16, "conv_umma_pack: weights must be fp32 or bf16");
    const Plan2 pl = make_plan2(Cin, Cout, k, stride);
    MGDT_CHECK(pl.ok, "conv_umma_pack: shape %d->%d k%d s%d is not supported by the tcgen05 path", Cin, Cout, k, stride);
    const long long total = (long long)pl.nsplit * pl.nks * pl.nmma_s * 2 * pl.Nc * 8;
    cudaStream_t s = (cudaStream_t)stream;
    const int g = cdiv(total, 256);
    if (w_dtype == MGDT_F32 && out_f16)
        launch_k(umma2_pack_kernel<float, __half>, dim3(g), dim3(256), 0, s, (const float*)w_ohwi, (__half*)packed, pl, Cin, Cout, k);
    else if (w_dtype == MGDT_F32)
        launch_k(umma2_pack_kernel<float, __nv_bfloat16>, dim3(g), dim3(256), 0, s, (const float*)w_ohwi, (__nv_bfloat16*)packed, pl, Cin, Cout, k);
    else if (out_f16)
        launch_k(umma2_pack_kernel<__nv_bfloat16, __half>, dim3(g), dim3(256), 0, s, (const __nv_bfloat16*)w_ohwi, (__half*)packed, pl, Cin, Cout, k);
    else
        launch_k(umma2_pack_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(g), dim3(256), 0, s, (const __nv_bfloat16*)w_ohwi, (__nv_bfloat16*)packed, pl, Cin, Cout, k);
    MGDT_LAUNCH_CHECK("umma_pack");
    return 0;
}

namespace mgdt {
// Per-image scaled copies of a packed bf16 weight image: out[img][chunk] = packed[chunk] * scale(img, column, ci .. ci + 7).
// One thread per 16-byte chunk (8 consecutive input channels of one (tap, plane, output column)).  Output columns are
// scaled in groups of `gcols`: group g < ngroups uses in_scale[g][img][ci], columns past the last group keep the
// weights unscaled (siblings that read the same map fused into ONE GEMM: the two TaskDecomposition convs, each with
// its own layer attention, and cls_prob_conv1 without any, nn/modules/head.py:509-521).
__global__ void umma2_scale_packed_kernel(const uint4* __restrict__ packed, uint4* __restrict__ out, Plan2 p, int Cin,
                                          const float* __restrict__ in_scale, unsigned per_img, unsigned total, int nimg,
                                          int ngroups, int gcols) {
    pdl_trigger();
    pdl_wait();
    const unsigned cps = (unsigned)p.nmma_s * 2u, Nc = (unsigned)p.Nc;
    for (unsigned c0 = blockIdx.x * blockDim.x + threadIdx.x; c0 < total; c0 += gridDim.x * blockDim.x) {
        const unsigned img = c0 / per_img, c = c0 - img * per_img;
        const unsigned chunk = (c / Nc) % cps, ks = (c / (Nc * cps)) % (unsigned)p.nks;
        const unsigned col = (c / (Nc * cps * (unsigned)p.nks)) * Nc + c % Nc;
        const int g = (int)(col / (unsigned)gcols);
        uint4 v = __ldg(packed + c);
        if (chunk < (unsigned)(p.taps * p.PS) && g < ngroups) {
            const unsigned plane = ks * (unsigned)p.PS + chunk % (unsigned)p.PS;
            const float4* sp = reinterpret_cast<const float4*>(in_scale + ((size_t)g * nimg + img) * Cin + plane * 8u);
            const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
            const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(w[j] << 16) * sc[2 * j],
                                                               __uint_as_float(w[j] & 0xffff0000u) * sc[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
        }
        out[c0] = v;
    }
}
}  // namespace mgdt

extern "C" int mgdt_conv_umma_pack_scaled_groups(const void* packed_bf16, int Cin, int Cout, int k, int stride, const float* in_scale,
                                                 int N, int ngroups, int group_cols, void* out, void* stream) {
    MGDT_CHECK(packed_bf16 && out && in_scale && N > 0 && ngroups > 0 && group_cols > 0, "conv_umma_pack_scaled: bad arguments");
    MGDT_CHECK((((uintptr_t)packed_bf16 | (uintptr_t)out | (uintptr_t)in_scale) & 15) == 0, "conv_umma_pack_scaled: pointers must be 16-byte aligned");
    MGDT_CHECK(k == 1 || ngroups == 1, "conv_umma_pack_scaled: column groups are for 1x1 layers");
    const Plan2 pl = make_plan2(Cin, Cout, k, stride);
    MGDT_CHECK(pl.ok, "conv_umma_pack_scaled: shape %d->%d k%d s%d is not supported by the tcgen05 path", Cin, Cout, k, stride);
    const long long per_img = (long long)pl.nsplit * pl.nks * pl.nmma_s * 2 * pl.Nc;   // 16-byte chunks
    MGDT_CHECK(per_img * N < (1LL << 31), "conv_umma_pack_scaled: too large");
    const int g = (int)std::min<long long>((per_img * N + 255) / 256, 148LL * 8);
    launch_k(umma2_scale_packed_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, (const uint4*)packed_bf16, (uint4*)out, pl, Cin, in_scale,
             (unsigned)per_img, (unsigned)(per_img * N), N, ngroups, group_cols);
    MGDT_LAUNCH_CHECK("umma_pack_scaled");
    return 0;
}

extern "C" int mgdt_conv_umma_pack_scaled(const void* packed_bf16, int Cin, int Cout, int k, int stride, const float* in_scale,
                                          int N, void* out, void* stream) {
    return mgdt_conv_umma_pack_scaled_groups(packed_bf16, Cin, Cout, k, stride, in_scale, N, 1, 1 << 30, out, stream);
}
