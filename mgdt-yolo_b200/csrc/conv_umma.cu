// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in TMEM).
//
// GEMM view: D[positions x Cout] = A[positions x K] * W[K x Cout], K = taps x Cin, pixels on M
// (UMMA M = 128), output channels on N.  One CTA owns MB blocks of 128 consecutive positions and
// Nc output channels; its whole K extent is resident in shared memory, so every input element is
// read from HBM/L2 once per CTA and the im2col expansion is never materialised:
//
//   * A lives in shared memory as channel planes  [Cin/8][parity][P positions][8 ch = 16 B], the
//     canonical no-swizzle K-major UMMA layout (a core matrix = 8 consecutive positions x 16 B).
//     For a 3x3 stride-1 conv the positions are linear indices into the zero-padded image
//     (row pitch W+2), so tap (dy,dx) is the SAME planes read through a descriptor whose start
//     address is shifted by ((dy-1)*(W+2) + (dx-1)) * 16 B ("shifted window"); the two junk columns
//     per row are computed and dropped in the epilogue.  Stride-2 convs split the padded input
//     into its 4 row/column parity sub-images (space-to-depth done while filling shared memory),
//     after which every tap is again a pure shift.
//   * W is pre-packed (mgdt_conv_umma_pack) into the matching K-major image [chunk][Nc][8].
//   * One elected thread issues all tcgen05.mma for the tile (K = 16 per instruction = two 16-byte
//     chunks; LBO addresses the second chunk, so chunk pairs may straddle taps), commits to an
//     mbarrier, and all 8 warps run the epilogue straight out of TMEM (tcgen05.ld 32x32b):
//     bias + activation + residual + bf16 pack + 16/32-byte stores into the NHWC destination slice.
//   * Input transforms of mgdt_conv2d (pre_add, in_scale, pix_scale, in_relu) are applied while
//     staging A.
//
// Overlap comes from co-residency (2 CTAs/SM: <=100 KB smem, <=256 TMEM columns each): one CTA
// fills shared memory while the other's MMAs / epilogue run.
#include "common.cuh"

#include <algorithm>

namespace mgdt {

constexpr int UM_THREADS = 256;
constexpr int UM_MAX_SMEM = 200 * 1024;

struct UmmaPlan {
    int mode;        // 0: 1x1 s1, 1: 3x3 s1, 2: 3x3 s2
    int planes;      // Cin / 8
    int npar;        // parity sub-images (1 or 4)
    int taps;        // 1 or 9
    int nchunks;     // taps * planes
    int nmma;        // ceil(nchunks / 2) K=16 instructions per (mb, tile)
    int Npad;        // Cout rounded up to 16
    int Nc;          // output channels per CTA (divides Npad, multiple of 16)
    int nsplit;      // Npad / Nc
    int tap_par[9], tap_dy[9], tap_dx[9];  // packed tap order (ascending smem address)
    bool ok;
};

struct UmmaRun {     // batch dependent part
    int MB, P, Wq, halo, pstride16, tiles_per_img, tmem_cols;
    long long tiles;
    size_t smem_a, smem_w, smem_total;
};

static UmmaPlan make_plan(int Cin, int Cout, int k, int stride) {
    UmmaPlan p{};
    p.ok = false;
    if (Cin % 8 != 0 || Cin < 8 || Cout < 1 || Cout > 512) return p;
    if (k == 1 && stride == 1) p.mode = 0;
    else if (k == 3 && stride == 1) p.mode = 1;
    else if (k == 3 && stride == 2) p.mode = 2;
    else return p;
    p.planes = Cin / 8;
    p.npar = p.mode == 2 ? 4 : 1;
    p.taps = p.mode == 0 ? 1 : 9;
    p.nchunks = p.taps * p.planes;
    p.nmma = (p.nchunks + 1) / 2;
    if ((p.planes & 1) && p.planes != 1 && p.taps > 1) return p;  // chunk pairs would straddle taps backwards
    p.Npad = (Cout + 15) / 16 * 16;
    // tap order: ascending (parity plane, shift) so that straddling pairs have a positive LBO
    int n = 0;
    if (p.mode == 0) {
        p.tap_par[0] = 0; p.tap_dy[0] = 0; p.tap_dx[0] = 0;
    } else if (p.mode == 1) {
        for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) { p.tap_par[n] = 0; p.tap_dy[n] = dy; p.tap_dx[n] = dx; ++n; }
    } else {
        for (int par = 0; par < 4; ++par)
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx)
                    if (((dy & 1) * 2 + (dx & 1)) == par) { p.tap_par[n] = par; p.tap_dy[n] = dy; p.tap_dx[n] = dx; ++n; }
    }
    // columns per CTA: whole Npad if the weight image stays under ~96 KB and N <= 256, else split
    const size_t per_col = (size_t)p.nmma * 2 * 16;  // bytes of weights per output channel
    int Nc = p.Npad;
    while ((Nc > 256 || per_col * Nc > 96 * 1024) && Nc % 32 == 0) Nc /= 2;
    if (Nc > 256 || per_col * Nc > 150 * 1024 || p.Npad % Nc != 0) return p;
    p.Nc = Nc;
    p.nsplit = p.Npad / Nc;
    p.ok = true;
    return p;
}

static bool try_run(const UmmaPlan& p, int MB, int N, int H, int W, int Ho, int Wo, UmmaRun& r) {
    if (MB * p.Nc > 512) return false;
    r.smem_w = (size_t)p.nmma * 2 * p.Nc * 16;
    r.MB = MB;
    if (p.mode == 0) {
        r.Wq = W; r.halo = 0; r.P = 128 * MB;
        r.tiles_per_img = 0;
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
    r.P = (r.P + 7) / 8 * 8 + 8;  // slack: the dummy chunk's LBO = 16 B read stays inside the buffer
    r.pstride16 = p.npar * r.P;
    if ((r.pstride16 & 1) == 0) r.pstride16 += 1;  // odd plane stride (16 B units): conflict-free staging stores
    r.smem_a = (size_t)p.planes * r.pstride16 * 16;
    r.smem_a = (r.smem_a + 127) / 128 * 128;
    r.smem_total = r.smem_a + r.smem_w + 256;
    if (r.smem_total > (size_t)UM_MAX_SMEM) return false;
    int cols = 32;
    while (cols < MB * p.Nc) cols <<= 1;
    r.tmem_cols = cols;
    return true;
}

// Largest tile (fewest halo re-reads) that still leaves every SM two CTAs; otherwise the smallest feasible.
static bool make_run(const UmmaPlan& p, int N, int H, int W, int Ho, int Wo, UmmaRun& r) {
    const int mbs[3] = {4, 2, 1};
    bool found = false;
    for (int i = 0; i < 3; ++i) {
        UmmaRun t;
        if (!try_run(p, mbs[i], N, H, W, Ho, Wo, t)) continue;
        r = t;
        found = true;
        if (t.tiles * p.nsplit >= 2 * 148) break;
    }
    return found;
}

// ---------------------------------------------------------------------------------- weight packing
// OHWI bf16 [Cout][k][k][Cin] -> [nsplit][2*nmma chunks][Nc][8] (zero padded)
__global__ void umma_pack_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out, UmmaPlan p,
                                 int Cin, int Cout, int k) {
    const long long total = (long long)p.nsplit * p.nmma * 2 * p.Nc * 8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % 8);
        const int nl = (int)((i / 8) % p.Nc);
        const int chunk = (int)((i / (8LL * p.Nc)) % (p.nmma * 2));
        const int ns = (int)(i / (8LL * p.Nc * p.nmma * 2));
        const int co = ns * p.Nc + nl;
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (chunk < p.nchunks && co < Cout) {
            const int t = chunk / p.planes, plane = chunk % p.planes;
            v = w[(((long long)co * k + p.tap_dy[t]) * k + p.tap_dx[t]) * Cin + plane * 8 + j];
        }
        out[i] = v;
    }
}

// ---------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // K-major, SWIZZLE_NONE ("interleave") shared-memory matrix descriptor, version 1 (Blackwell)
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

struct UmmaP {
    const __nv_bfloat16 *x, *w, *pre_add, *pix_scale, *residual;
    const float *bias, *in_scale;
    __nv_bfloat16* y;
    int N, H, W, Cin, Cout, Ho, Wo;
    int x_cs, y_cs, add_cs, ps_cs, res_cs, act, in_relu;
    int y_vec, res_vec;  // 16-byte aligned destination / residual rows
    UmmaPlan pl;
    UmmaRun rn;
    long long M_total;   // mode 0: N*H*W
    // DCNv2 staging (mode 0 over a virtual 9*dcn_cin-channel input built by modulated bilinear sampling)
    const __nv_bfloat16 *dcn_off, *dcn_mask;
    int off_cs, mask_cs, mask_logit, dcn_cin;
};

// position (tile-relative index m in [0, 128*MB)) -> output pixel index (n*Ho*Wo + ho*Wo + wo) or -1
__device__ __forceinline__ long long out_pixel(const UmmaP& p, long long tile, int m, int& n_img) {
    if (p.pl.mode == 0) {
        const long long pix = tile * (128LL * p.rn.MB) + m;
        if (pix >= p.M_total) return -1;
        n_img = (int)(pix / ((long long)p.H * p.W));
        return pix;
    }
    const int n = (int)(tile / p.rn.tiles_per_img);
    const int tt = (int)(tile - (long long)n * p.rn.tiles_per_img);
    n_img = n;
    if (p.pl.mode == 1) {
        const int q = p.rn.Wq + tt * 128 * p.rn.MB + m;
        const int hp = q / p.rn.Wq, wp = q - hp * p.rn.Wq;
        if (hp < 1 || hp > p.H || wp < 1 || wp > p.W) return -1;
        return ((long long)n * p.H + (hp - 1)) * p.W + (wp - 1);
    }
    const int q = tt * 128 * p.rn.MB + m;
    const int ho = q / p.rn.Wq, wo = q - ho * p.rn.Wq;
    if (ho >= p.Ho || wo >= p.Wo) return -1;
    return ((long long)n * p.Ho + ho) * p.Wo + wo;
}

__global__ void __launch_bounds__(UM_THREADS, 1) conv_umma_kernel(const UmmaP p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tile = blockIdx.x;
    const int ns = blockIdx.y;
    const UmmaPlan& pl = p.pl;
    const UmmaRun& rn = p.rn;

    unsigned char* sA = smem;
    unsigned char* sW = smem + rn.smem_a;
    const uint32_t bar = smem_u32(&mbar);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"((uint32_t)rn.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // ---------------- stage W: contiguous packed image of this column split
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.w) + (size_t)ns * (rn.smem_w / 16);
        uint4* dst = reinterpret_cast<uint4*>(sW);
        const int n16 = (int)(rn.smem_w / 16);
        for (int i = tid; i < n16; i += UM_THREADS) dst[i] = __ldg(src + i);
    }

    // ---------------- stage A: [plane][parity][P][16 B]
    {
        const int planes = pl.planes, P = rn.P, npar = pl.npar;
        const int total = planes * P * npar;
        int n_img = 0, tt = 0;
        long long pix0 = 0;
        if (pl.mode == 0) {
            pix0 = tile * (128LL * rn.MB);
        } else {
            n_img = (int)(tile / rn.tiles_per_img);
            tt = (int)(tile - (long long)n_img * rn.tiles_per_img);
        }
        if (p.dcn_off) {
            // DCNv2 (DyDCNv2.forward, block.py:427-429): chunk (pos, plane) = 8 channels of tap = plane / (Cin/8),
            // value = mask * bilinear(x, (h + ky - 1 + dy, w + kx - 1 + dx)), zero outside (-1,H)x(-1,W).
            const int cgs = p.dcn_cin / 8;
            for (int e = tid; e < total; e += UM_THREADS) {
                const int plane = e % planes;
                const int pos = e / planes;
                const long long g = pix0 + pos;
                float f[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.f;
                if (g < p.M_total) {
                    const int tap = plane / cgs, cg = plane - tap * cgs;
                    const int n = (int)(g / ((long long)p.H * p.W));
                    const int rem = (int)(g - (long long)n * p.H * p.W);
                    const int hq = rem / p.W, wq = rem - hq * p.W;
                    const __nv_bfloat16* ofp = p.dcn_off + g * p.off_cs + 2 * tap;
                    const float dy = __bfloat162float(ofp[0]), dx = __bfloat162float(ofp[1]);
                    float m = __bfloat162float(p.dcn_mask[g * p.mask_cs + tap]);
                    if (p.mask_logit) m = sigmoidf_(m);
                    const float py = (float)(hq + tap / 3 - 1) + dy, px = (float)(wq + tap % 3 - 1) + dx;
                    if (py > -1.f && py < (float)p.H && px > -1.f && px < (float)p.W) {
                        const int y0 = (int)floorf(py), x0 = (int)floorf(px);
                        const float ly = py - (float)y0, lx = px - (float)x0;
                        const float hy = 1.f - ly, hx = 1.f - lx;
                        const __nv_bfloat16* xn = p.x + (long long)n * p.H * p.W * p.x_cs + cg * 8;
                        const float wgt[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
                        const int yy[4] = {y0, y0, y0 + 1, y0 + 1}, xx[4] = {x0, x0 + 1, x0, x0 + 1};
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            if (yy[c4] >= 0 && yy[c4] <= p.H - 1 && xx[c4] >= 0 && xx[c4] <= p.W - 1) {
                                const uint4 v = __ldg(reinterpret_cast<const uint4*>(xn + ((long long)yy[c4] * p.W + xx[c4]) * p.x_cs));
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
                *reinterpret_cast<uint4*>(sA + ((size_t)plane * rn.pstride16 + pos) * 16) = o;
            }
        } else
        for (int e = tid; e < total; e += UM_THREADS) {
            const int plane = e % planes;
            const int rest = e / planes;
            const int pos = rest % P;
            const int par = rest / P;
            long long pix = -1;  // input pixel index n*H*W + h*W + w
            int n = n_img;
            if (pl.mode == 0) {
                const long long g = pix0 + pos;
                if (g < p.M_total) { pix = g; n = (int)(g / ((long long)p.H * p.W)); }
            } else if (pl.mode == 1) {
                const int q = rn.Wq + tt * 128 * rn.MB - rn.halo + pos;   // padded linear index
                const int hp = q / rn.Wq, wp = q - hp * rn.Wq;
                if (q >= 0 && hp >= 1 && hp <= p.H && wp >= 1 && wp <= p.W)
                    pix = ((long long)n * p.H + (hp - 1)) * p.W + (wp - 1);
            } else {
                const int q = tt * 128 * rn.MB + pos;                      // index into the parity sub-image
                const int r = q / rn.Wq, c = q - r * rn.Wq;
                const int hi = 2 * r + (par >> 1) - 1, wi = 2 * c + (par & 1) - 1;
                if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) pix = ((long long)n * p.H + hi) * p.W + wi;
            }
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (pix >= 0) {
                v = __ldg(reinterpret_cast<const uint4*>(p.x + pix * p.x_cs + plane * 8));
                if (p.pre_add || p.in_scale || p.pix_scale || p.in_relu) {
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
                    if (p.pre_add) {
                        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.pre_add + pix * p.add_cs + plane * 8));
                        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(ah[j]); f[2 * j] += t.x; f[2 * j + 1] += t.y; }
                    }
                    if (p.in_scale) {
                        const float* s = p.in_scale + (long long)n * p.Cin + plane * 8;
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] *= s[j];
                    }
                    if (p.pix_scale) {
                        const float s = __bfloat162float(p.pix_scale[pix * p.ps_cs]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] *= s;
                    }
                    if (p.in_relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                }
            }
            *reinterpret_cast<uint4*>(sA + ((size_t)plane * rn.pstride16 + (size_t)par * P + pos) * 16) = v;
        }
    }
    // generic-proxy writes -> visible to the tensor core (async proxy), then CTA-wide sync
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    // ---------------- MMA issue (one thread)
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(pl.Nc >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
        const uint32_t b_lbo = (uint32_t)pl.Nc * 16;
        for (int mb = 0; mb < rn.MB; ++mb) {
            for (int i = 0; i < pl.nmma; ++i) {
                const int c0 = 2 * i, c1 = 2 * i + 1;
                // byte offset of a chunk's window inside sA (for block mb = 0)
                auto off = [&](int c) -> uint32_t {
                    const int t = c / pl.planes, plane = c - t * pl.planes;
                    int shift;
                    if (pl.mode == 0) shift = 0;
                    else if (pl.mode == 1) shift = rn.halo + (pl.tap_dy[t] - 1) * rn.Wq + (pl.tap_dx[t] - 1);
                    else shift = (pl.tap_dy[t] >> 1) * rn.Wq + (pl.tap_dx[t] >> 1);
                    return (uint32_t)(((size_t)plane * rn.pstride16 + (size_t)pl.tap_par[t] * rn.P + shift) * 16);
                };
                const uint32_t o0 = off(c0);
                const uint32_t lbo = (c1 < pl.nchunks) ? (off(c1) - o0) : 16u;  // dummy chunk: weights are zero
                const uint64_t adesc = make_desc(a0 + o0 + (uint32_t)mb * 2048u, lbo, 128u);
                const uint64_t bdesc = make_desc(w0 + (uint32_t)c0 * b_lbo, b_lbo, 128u);
                const uint32_t d = tmem_base + (uint32_t)(mb * pl.Nc);
                const uint32_t acc = i > 0 ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                    ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }

    // ---------------- wait for the accumulators (bounded spin: a lost commit traps instead of hanging)
    {
        const long long t0 = clock64();
        while (!mbar_try_wait(bar, 0)) {
            if (clock64() - t0 > 4000000000LL) __trap();
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---------------- epilogue: TMEM -> registers -> bias/act/residual -> NHWC bf16
    {
        const int quad = warp & 3, half = warp >> 2;
        const int ncch = pl.Nc / 16;
        for (int mb = 0; mb < rn.MB; ++mb) {
            const int m = mb * 128 + quad * 32 + lane;
            int n_img;
            const long long opix = out_pixel(p, tile, m, n_img);
            for (int cc = half; cc < ncch; cc += 2) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(mb * pl.Nc + cc * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (opix < 0) continue;
                const int co0 = ns * pl.Nc + cc * 16;
                if (co0 >= p.Cout) continue;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float t = __uint_as_float(r[j]);
                    const int co = co0 + j;
                    if (p.bias && co < p.Cout) t += p.bias[co];
                    v[j] = apply_act(t, p.act);
                }
                __nv_bfloat16* yp = p.y + opix * p.y_cs + co0;
                const bool full = co0 + 16 <= p.Cout;
                if (p.residual) {
                    const __nv_bfloat16* rp = p.residual + opix * p.res_cs + co0;
                    if (full && p.res_vec) {
                        const uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
                        const uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                        const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
                        const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 ta = __bfloat1622float2(ah[j]), tb = __bfloat1622float2(bh[j]);
                            v[2 * j] += ta.x; v[2 * j + 1] += ta.y; v[8 + 2 * j] += tb.x; v[8 + 2 * j + 1] += tb.y;
                        }
                    } else {
                        for (int j = 0; j < 16 && co0 + j < p.Cout; ++j) v[j] += __bfloat162float(rp[j]);
                    }
                }
                if (full && p.y_vec) {
                    uint4 o0, o1;
                    __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
                    __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
                    }
                    reinterpret_cast<uint4*>(yp)[0] = o0;
                    reinterpret_cast<uint4*>(yp)[1] = o1;
                } else {
                    for (int j = 0; j < 16 && co0 + j < p.Cout; ++j) yp[j] = __float2bfloat16_rn(v[j]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)rn.tmem_cols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------- host side
static bool plan_for(const mgdt_conv_args* a, UmmaPlan& pl, UmmaRun& rn, int& Ho, int& Wo) {
    if (a->dtype != MGDT_BF16 || a->kh != a->kw) return false;
    if (a->pad != a->kh / 2) return false;
    pl = make_plan(a->Cin, a->Cout, a->kh, a->stride);
    if (!pl.ok) return false;
    Ho = (a->H + 2 * a->pad - a->kh) / a->stride + 1;
    Wo = (a->W + 2 * a->pad - a->kw) / a->stride + 1;
    if (!make_run(pl, a->N, a->H, a->W, Ho, Wo, rn)) return false;
    if (rn.tiles > 0x7fffffffLL) return false;
    return true;
}

bool conv2d_umma_supported(const mgdt_conv_args* a) {
    if (!a->w_umma) return false;
    UmmaPlan pl; UmmaRun rn; int Ho, Wo;
    if (!plan_for(a, pl, rn, Ho, Wo)) return false;
    // 16-byte loads of 8-channel chunks
    if (((uintptr_t)a->x & 15) || (a->x_cs & 7)) return false;
    if (a->pre_add && (((uintptr_t)a->pre_add & 15) || (a->add_cs & 7))) return false;
    if ((uintptr_t)a->w_umma & 15) return false;
    return true;
}

int conv2d_umma(const mgdt_conv_args* a, cudaStream_t s) {
    UmmaPlan pl; UmmaRun rn; int Ho, Wo;
    if (!plan_for(a, pl, rn, Ho, Wo)) return set_error(-EINVAL, "conv2d_umma: unsupported shape");
    UmmaP p;
    p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w_umma;
    p.pre_add = (const __nv_bfloat16*)a->pre_add; p.pix_scale = (const __nv_bfloat16*)a->pix_scale;
    p.residual = (const __nv_bfloat16*)a->residual; p.bias = a->bias; p.in_scale = a->in_scale;
    p.y = (__nv_bfloat16*)a->y;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout; p.Ho = Ho; p.Wo = Wo;
    p.x_cs = a->x_cs; p.y_cs = a->y_cs; p.add_cs = a->add_cs; p.ps_cs = a->ps_cs; p.res_cs = a->res_cs;
    p.act = a->act; p.in_relu = a->in_relu;
    p.y_vec = (((uintptr_t)a->y & 15) == 0 && (a->y_cs & 7) == 0) ? 1 : 0;
    p.res_vec = (a->residual && ((uintptr_t)a->residual & 15) == 0 && (a->res_cs & 7) == 0) ? 1 : 0;
    p.pl = pl; p.rn = rn;
    p.M_total = (long long)a->N * a->H * a->W;
    p.dcn_off = nullptr; p.dcn_mask = nullptr; p.off_cs = p.mask_cs = p.mask_logit = p.dcn_cin = 0;
    {
        cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_MAX_SMEM);
        if (e != cudaSuccess) return set_error(-EIO, "conv2d_umma: smem attr: %s", cudaGetErrorString(e));
    }
    dim3 grid((unsigned)rn.tiles, (unsigned)pl.nsplit);
    conv_umma_kernel<<<grid, UM_THREADS, rn.smem_total, s>>>(p);
    MGDT_LAUNCH_CHECK("conv_umma");
    return 0;
}

// DCNv2 3x3 on the tensor cores: a 1x1 GEMM over the virtual (9*Cin)-channel im2col built while staging.
bool dcn_umma_supported(const void* x, int x_cs, const void* w_umma, int N, int H, int W, int Cin, int Cout) {
    if (!w_umma || Cin % 8 != 0 || ((uintptr_t)x & 15) || (x_cs & 7) || ((uintptr_t)w_umma & 15)) return false;
    UmmaPlan pl = make_plan(9 * Cin, Cout, 1, 1);
    UmmaRun rn;
    return pl.ok && make_run(pl, N, H, W, H, W, rn) && rn.tiles <= 0x7fffffffLL;
}

int dcn_umma(const void* x, int x_cs, const void* offset, int off_cs, const void* mask, int mask_cs, int mask_is_logit,
             const void* w_umma, void* y, int y_cs, int N, int H, int W, int Cin, int Cout, cudaStream_t s) {
    UmmaPlan pl = make_plan(9 * Cin, Cout, 1, 1);
    UmmaRun rn;
    if (!pl.ok || !make_run(pl, N, H, W, H, W, rn)) return set_error(-EINVAL, "dcn_umma: unsupported shape");
    UmmaP p;
    p.x = (const __nv_bfloat16*)x; p.w = (const __nv_bfloat16*)w_umma; p.pre_add = nullptr; p.pix_scale = nullptr;
    p.residual = nullptr; p.bias = nullptr; p.in_scale = nullptr; p.y = (__nv_bfloat16*)y;
    p.N = N; p.H = H; p.W = W; p.Cin = 9 * Cin; p.Cout = Cout; p.Ho = H; p.Wo = W;
    p.x_cs = x_cs; p.y_cs = y_cs; p.add_cs = p.ps_cs = p.res_cs = 0; p.act = MGDT_ACT_NONE; p.in_relu = 0;
    p.y_vec = (((uintptr_t)y & 15) == 0 && (y_cs & 7) == 0) ? 1 : 0;
    p.res_vec = 0;
    p.pl = pl; p.rn = rn;
    p.M_total = (long long)N * H * W;
    p.dcn_off = (const __nv_bfloat16*)offset; p.dcn_mask = (const __nv_bfloat16*)mask;
    p.off_cs = off_cs; p.mask_cs = mask_cs; p.mask_logit = mask_is_logit; p.dcn_cin = Cin;
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_MAX_SMEM);
    if (e != cudaSuccess) return set_error(-EIO, "dcn_umma: smem attr: %s", cudaGetErrorString(e));
    conv_umma_kernel<<<dim3((unsigned)rn.tiles, (unsigned)pl.nsplit), UM_THREADS, rn.smem_total, s>>>(p);
    MGDT_LAUNCH_CHECK("dcn_umma");
    return 0;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" size_t mgdt_conv_umma_packed_bytes(int Cin, int Cout, int k, int stride) {
    const UmmaPlan pl = make_plan(Cin, Cout, k, stride);
    if (!pl.ok) return 0;
    return (size_t)pl.nsplit * pl.nmma * 2 * pl.Nc * 16;
}

extern "C" int mgdt_conv_umma_pack(const void* w_ohwi, int Cin, int Cout, int k, int stride, void* packed, void* stream) {
    MGDT_CHECK(w_ohwi && packed, "conv_umma_pack: null pointer");
    const UmmaPlan pl = make_plan(Cin, Cout, k, stride);
    MGDT_CHECK(pl.ok, "conv_umma_pack: shape %d->%d k%d s%d is not supported by the tcgen05 path", Cin, Cout, k, stride);
    const long long total = (long long)pl.nsplit * pl.nmma * 2 * pl.Nc * 8;
    umma_pack_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)w_ohwi,
                                                                        (__nv_bfloat16*)packed, pl, Cin, Cout, k);
    MGDT_LAUNCH_CHECK("umma_pack");
    return 0;
}
