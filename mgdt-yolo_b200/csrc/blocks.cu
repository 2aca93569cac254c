// Block-specific kernels: ConvNeXtV2 depthwise 7x7 + LayerNorm, modulated deformable conv
// (DCNv2), and the DFL / dist2bbox decode.
#include "common.cuh"
#include <string.h>

namespace mgdt {

// ------------------------------------------------------------------ dwconv7 + LN
// One warp per output pixel; lanes stride over channels (NHWC: contiguous), so every tap is a
// coalesced row read and the LayerNorm reduction is a warp shuffle.  Up to 8 channels per lane
// (C <= 256) are kept in registers.
constexpr int DW_WARPS = 8;
constexpr int DW_MAXPL = 8;

template <typename T>
__global__ void __launch_bounds__(DW_WARPS * 32) dwconv7_ln_kernel(const T* __restrict__ x, int x_cs,
                                                                   const T* __restrict__ w, const float* __restrict__ bias,
                                                                   const float* __restrict__ ln_w,
                                                                   const float* __restrict__ ln_b, float eps,
                                                                   T* __restrict__ y, int y_cs, int H, int W, int C,
                                                                   long long npix) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const long long pix = (long long)blockIdx.x * DW_WARPS + (threadIdx.x >> 5);
    if (pix >= npix) return;
    const int wq = (int)(pix % W);
    const int hq = (int)((pix / W) % H);
    const long long n = pix / ((long long)W * H);
    const T* xn = x + n * (long long)H * W * x_cs;
    float acc[DW_MAXPL];
#pragma unroll
    for (int j = 0; j < DW_MAXPL; ++j) {
        const int c = lane + 32 * j;
        acc[j] = (c < C) ? bias[c] : 0.f;
    }
    for (int dy = 0; dy < 7; ++dy) {
        const int hh = hq + dy - 3;
        if (hh < 0 || hh >= H) continue;
        for (int dx = 0; dx < 7; ++dx) {
            const int ww = wq + dx - 3;
            if (ww < 0 || ww >= W) continue;
            const T* xp = xn + ((long long)hh * W + ww) * x_cs;
            const T* wp = w + (dy * 7 + dx) * C;
#pragma unroll
            for (int j = 0; j < DW_MAXPL; ++j) {
                const int c = lane + 32 * j;
                if (c < C) acc[j] = fmaf(ldf(xp + c), ldf(wp + c), acc[j]);
            }
        }
    }
    // LayerNorm over C (biased variance, two-pass in registers)
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < DW_MAXPL; ++j)
        if (lane + 32 * j < C) s += acc[j];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < DW_MAXPL; ++j)
        if (lane + 32 * j < C) {
            const float d = acc[j] - mean;
            q += d * d;
        }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    T* yp = y + pix * y_cs;
#pragma unroll
    for (int j = 0; j < DW_MAXPL; ++j) {
        const int c = lane + 32 * j;
        if (c < C) stf(yp + c, (acc[j] - mean) * rstd * ln_w[c] + ln_b[c]);
    }
}

// ------------------------------------------------------------------ dwconv7 + LN, shared-memory tiled
// CTA = 8x8 output pixels of one image, all C channels: the (8+6)x(8+6)xC input halo tile and the 49xC
// weights live in shared memory; warp r computes output row r (8 pixels, lanes over channels, CPL
// channels per lane) re-using every staged input value for the up-to-7 outputs it contributes to, then
// LayerNorm is a warp reduction per pixel.  ~200 instructions per pixel-lane instead of ~500 global loads.
static int g_dw_pairs = 1;       // option "dw_pairs": channel-pair kernel for bf16 (dwconv7_ln_pairs)
int blocks_set_option(const char* name, int value) {
    if (!strcmp(name, "dw_pairs")) { g_dw_pairs = value ? 1 : 0; return 1; }
    return 0;
}
constexpr int DT = 8;            // tile edge
constexpr int DTI = DT + 6;      // input tile edge

template <typename T, int CPL, bool VEC>
__global__ void __launch_bounds__(256) dwconv7_ln_tiled(const T* __restrict__ x, int x_cs, const T* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, T* __restrict__ y,
                                                        int y_cs, int H, int W, int C, int tiles_x, int tiles_y) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char dsm[];
    const int CP = CPL * 32;                                   // padded channel count
    float* sw = reinterpret_cast<float*>(dsm);                 // [49][CP]
    float* sx = sw + 49 * CP;                                  // [DTI][DTI][CP], staged as fp32
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = blockIdx.x;
    const int tx = b % tiles_x; b /= tiles_x;
    const int ty = b % tiles_y;
    const int n = b / tiles_y;
    const int h0 = ty * DT, w0 = tx * DT;
    const T* xn = x + (size_t)n * H * W * x_cs;
    if (VEC) {
        // bf16, C % 8 == 0, 16-byte aligned rows.  The input halo tile stays bf16 in shared memory and is staged by
        // cp.async (zero-filled outside the image / beyond C) without a register round trip: half the footprint of an
        // fp32 tile, so four CTAs share an SM and the staging of one overlaps the arithmetic of the others.  The
        // weights are converted to fp32 once while the copies are in flight.
        constexpr int C8 = CPL * 4;   // 8-channel chunks per padded pixel
        {
            const uint32_t sx32 = (uint32_t)__cvta_generic_to_shared(sx);
            const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(xn);
            for (int i = tid; i < DTI * DTI * C8; i += 256) {
                const int pix = i / C8, c8 = (i - pix * C8) * 8;
                const int iy = pix / DTI, ix = pix - iy * DTI;
                const int hh = h0 + iy - 3, ww = w0 + ix - 3;
                const bool ok = c8 < C && hh >= 0 && hh < H && ww >= 0 && ww < W;
                const __nv_bfloat16* src = ok ? xb + (size_t)(hh * W + ww) * x_cs + c8 : xb;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sx32 + (uint32_t)(pix * CP + c8) * 2u), "l"(src),
                             "r"(ok ? 16u : 0u) : "memory");
            }
        }
        for (int i = tid; i < 49 * C8; i += 256) {
            const int tap = i / C8, c8 = (i - tap * C8) * 8;
            float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
            if (c8 < C) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(w) + tap * C + c8));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
                const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
                lo = make_float4(a.x, a.y, b.x, b.y); hi = make_float4(c.x, c.y, d.x, d.y);
            }
            float4* dst = reinterpret_cast<float4*>(sw + tap * CP + c8);
            dst[0] = lo; dst[1] = hi;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
        for (int i = tid; i < 49 * CP; i += 256) {
            const int c = i % CP;
            sw[i] = c < C ? ldf(w + (i / CP) * C + c) : 0.f;
        }
        for (int i = tid; i < DTI * DTI * CP; i += 256) {
            const int c = i % CP, pix = i / CP;
            const int iy = pix / DTI, ix = pix - iy * DTI;
            const int hh = h0 + iy - 3, ww = w0 + ix - 3;
            float v = 0.f;
            if (c < C && hh >= 0 && hh < H && ww >= 0 && ww < W) v = ldf(xn + (size_t)(hh * W + ww) * x_cs + c);
            sx[i] = v;
        }
    }
    __syncthreads();
    const int r = warp;  // output row inside the tile
    float acc[DT][CPL];
#pragma unroll
    for (int px = 0; px < DT; ++px)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int c = lane + 32 * j;
            acc[px][j] = c < C ? bias[c] : 0.f;
        }
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
        float wk[7][CPL];
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int j = 0; j < CPL; ++j) wk[dx][j] = sw[(dy * 7 + dx) * CP + lane + 32 * j];
        const float* row = sx + (size_t)((r + dy) * DTI) * CP;
        const unsigned short* rowh = reinterpret_cast<const unsigned short*>(sx) + (size_t)((r + dy) * DTI) * CP;   // VEC: bf16 tile
#pragma unroll
        for (int ix = 0; ix < DTI; ++ix) {
            float v[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                if (VEC) v[j] = __uint_as_float((uint32_t)rowh[ix * CP + lane + 32 * j] << 16);
                else v[j] = row[ix * CP + lane + 32 * j];
            }
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const int px = ix - dx;
                if (px >= 0 && px < DT) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) acc[px][j] = fmaf(v[j], wk[dx][j], acc[px][j]);
                }
            }
        }
    }
    const int hq = h0 + r;
    float lw[CPL], lb[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        lw[j] = c < C ? ln_w[c] : 0.f;
        lb[j] = c < C ? ln_b[c] : 0.f;
    }
#pragma unroll
    for (int px = 0; px < DT; ++px) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (lane + 32 * j < C) s += acc[px][j];
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (lane + 32 * j < C) {
                const float d = acc[px][j] - mean;
                q += d * d;
            }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        const int wq = w0 + px;
        if (hq < H && wq < W) {
            T* yp = y + ((size_t)n * H * W + (size_t)hq * W + wq) * y_cs;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int c = lane + 32 * j;
                if (c < C) stf(yp + c, (acc[px][j] - mean) * rstd * lw[j] + lb[j]);
            }
        }
    }
}

// Channel-PAIR form of the tiled kernel (bf16, C % 8 == 0): ncu showed dwconv7_ln_tiled issue-bound (IPC 2.3, "not selected"
// 30 %) at ~3,500 instructions per 8-pixel row -- one LDS.U16 + shift + FFMA per channel and tap.  Here a lane owns channel
// pairs (2q, 2q + 1), q = l16 + 16 j: one LDS.32 fetches both values, two ALU instructions unpack them and one packed
// FFMA2 (fma.rn.f32x2, bit-identical to two FFMAs) does both channels.  A half-warp holds all C channels of a pixel, so a
// warp computes TWO output rows x four pixels; weights stay bf16 pairs in shared memory (48 KB per CTA: four CTAs / SM).
// The tile's row pitch is padded by 16 words so that the two half-warps (rows r, r + 1) read disjoint banks.
// Same accumulation order per output as dwconv7_ln_tiled (bias, then dy, ix ascending).
template <int CPL>
__global__ void __launch_bounds__(256) dwconv7_ln_pairs(const __nv_bfloat16* __restrict__ x, int x_cs, const __nv_bfloat16* __restrict__ w,
                                                        const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                        const float* __restrict__ ln_b, float eps, __nv_bfloat16* __restrict__ y,
                                                        int y_cs, int H, int W, int C, int tiles_x, int tiles_y) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char dsm[];
    constexpr int CP = CPL * 32, CW = CP / 2;                    // padded channels, 32-bit words per pixel
    constexpr int C8 = CPL * 4;                                  // 8-channel chunks per padded pixel
    constexpr int ROWW = DTI * CW + 16;                          // words per tile row (padded)
    uint32_t* sw = reinterpret_cast<uint32_t*>(dsm);             // [49][CW] bf16 pairs
    uint32_t* sx = sw + 49 * CW;                                 // [DTI][ROWW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int b = blockIdx.x;
    const int tx = b % tiles_x; b /= tiles_x;
    const int ty = b % tiles_y;
    const int n = b / tiles_y;
    const int h0 = ty * DT, w0 = tx * DT;
    const __nv_bfloat16* xn = x + (size_t)n * H * W * x_cs;
    {
        const uint32_t sx32 = (uint32_t)__cvta_generic_to_shared(sx), sw32 = (uint32_t)__cvta_generic_to_shared(sw);
        // (ncu: the index arithmetic of a flat chunk loop -- two divisions, 64-bit addressing per 16-byte chunk -- was half
        // of the kernel's instructions.)  A thread keeps ONE (column, channel chunk) of the tile and walks its 14 rows:
        // everything but the row test is computed once.
        for (int t = tid; t < DTI * C8; t += 256) {
            const int ix = t / C8, c8 = (t - ix * C8) * 8;
            const int ww = w0 + ix - 3;
            const bool colok = c8 < C && ww >= 0 && ww < W;
            const __nv_bfloat16* src = xn + (ptrdiff_t)((h0 - 3) * W + ww) * x_cs + c8;
            uint32_t dst = sx32 + (uint32_t)(ix * CW) * 4u + (uint32_t)c8 * 2u;
            const ptrdiff_t rstep = (ptrdiff_t)W * x_cs;
#pragma unroll
            for (int iy = 0; iy < DTI; ++iy, src += rstep, dst += ROWW * 4u) {
                const bool ok = colok && (unsigned)(h0 + iy - 3) < (unsigned)H;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? src : xn), "r"(ok ? 16u : 0u) : "memory");
            }
        }
        for (int i = tid; i < 49 * C8; i += 256) {
            const int tap = i / C8, c8 = (i - tap * C8) * 8;
            const bool ok = c8 < C;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sw32 + (uint32_t)(tap * CW) * 4u + (uint32_t)c8 * 2u),
                         "l"(ok ? w + tap * C + c8 : w), "r"(ok ? 16u : 0u) : "memory");
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();
    const int half = lane >> 4, l16 = lane & 15;
    const int r = 2 * (warp >> 1) + half;          // output row inside the tile
    const int px0 = 4 * (warp & 1);                // first of this warp's four output pixels
    float2 acc[4][CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = 2 * (l16 + 16 * j);
        const float2 bj = c < C ? make_float2(bias[c], bias[c + 1]) : make_float2(0.f, 0.f);
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[px][j] = bj;
    }
    auto unpack = [](uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); };
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
        float2 wk[7][CPL];
#pragma unroll
        for (int dx = 0; dx < 7; ++dx)
#pragma unroll
            for (int j = 0; j < CPL; ++j) wk[dx][j] = unpack(sw[(dy * 7 + dx) * CW + l16 + 16 * j]);
        const uint32_t* row = sx + (size_t)(r + dy) * ROWW + px0 * CW;
#pragma unroll
        for (int ixl = 0; ixl < 10; ++ixl) {       // input columns px0 .. px0 + 9 of the tile
            float2 v[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) v[j] = unpack(row[ixl * CW + l16 + 16 * j]);
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const int px = ixl - dx;
                if (px >= 0 && px < 4) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) acc[px][j] = __ffma2_rn(v[j], wk[dx][j], acc[px][j]);
                }
            }
        }
    }
    const int hq = h0 + r;
    float2 lw[CPL], lb[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const int c = 2 * (l16 + 16 * j);
        lw[j] = c < C ? make_float2(ln_w[c], ln_w[c + 1]) : make_float2(0.f, 0.f);
        lb[j] = c < C ? make_float2(ln_b[c], ln_b[c + 1]) : make_float2(0.f, 0.f);
    }
    auto half_sum = [](float v) {                  // over the 16 lanes of this half-warp
#pragma unroll
        for (int o = 8; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (2 * (l16 + 16 * j) < C) s += acc[px][j].x + acc[px][j].y;
        const float mean = half_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
            if (2 * (l16 + 16 * j) < C) {
                const float d0 = acc[px][j].x - mean, d1 = acc[px][j].y - mean;
                q += d0 * d0 + d1 * d1;
            }
        const float rstd = rsqrtf(half_sum(q) / (float)C + eps);
        const int wq = w0 + px0 + px;
        if (hq < H && wq < W) {
            __nv_bfloat16* yp = y + ((size_t)n * H * W + (size_t)hq * W + wq) * y_cs;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int c = 2 * (l16 + 16 * j);
                if (c < C)
                    *reinterpret_cast<__nv_bfloat162*>(yp + c) = __floats2bfloat162_rn((acc[px][j].x - mean) * rstd * lw[j].x + lb[j].x,
                                                                                       (acc[px][j].y - mean) * rstd * lw[j].y + lb[j].y);
            }
        }
    }
}

template <typename T, int CPL>
static int launch_dw_tiled(const void* x, int x_cs, const void* w, const float* bias, const float* ln_w,
                           const float* ln_b, float eps, void* y, int y_cs, int N, int H, int W, int C, cudaStream_t s) {
    const int CP = CPL * 32;
    const int tx = (W + DT - 1) / DT, ty = (H + DT - 1) / DT;
    const bool vec = sizeof(T) == 2 && C % 8 == 0 && x_cs % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0;
    const size_t smem = sizeof(float) * 49 * CP + (vec ? 2 : sizeof(float)) * DTI * DTI * CP;
    cudaError_t e;
    if constexpr (sizeof(T) == 2) {
        if (vec && g_dw_pairs && (y_cs & 1) == 0 && ((uintptr_t)y & 3) == 0) {
            const size_t smem2 = 4 * (size_t)(49 * (CP / 2) + DTI * (DTI * (CP / 2) + 16));
            e = cudaFuncSetAttribute(dwconv7_ln_pairs<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return set_error(-EIO, "dwconv7_ln: smem attr: %s", cudaGetErrorString(e));
            launch_k(dwconv7_ln_pairs<CPL>, dim3(tx * ty * N), dim3(256), smem2, s, (const __nv_bfloat16*)x, x_cs, (const __nv_bfloat16*)w, bias,
                     ln_w, ln_b, eps, (__nv_bfloat16*)y, y_cs, H, W, C, tx, ty);
            return 0;
        }
    }
    if (vec) {
        e = cudaFuncSetAttribute(dwconv7_ln_tiled<T, CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-EIO, "dwconv7_ln: smem attr: %s", cudaGetErrorString(e));
        launch_k(dwconv7_ln_tiled<T, CPL, true>, dim3(tx * ty * N), dim3(256), smem, s, (const T*)x, x_cs, (const T*)w, bias, ln_w, ln_b, eps,
                                                                      (T*)y, y_cs, H, W, C, tx, ty);
    } else {
        e = cudaFuncSetAttribute(dwconv7_ln_tiled<T, CPL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-EIO, "dwconv7_ln: smem attr: %s", cudaGetErrorString(e));
        launch_k(dwconv7_ln_tiled<T, CPL, false>, dim3(tx * ty * N), dim3(256), smem, s, (const T*)x, x_cs, (const T*)w, bias, ln_w, ln_b, eps,
                                                                       (T*)y, y_cs, H, W, C, tx, ty);
    }
    return 0;
}

// ------------------------------------------------------------------ DCNv2 3x3
// CTA = DCN_PIX output pixels.  Phase 1 builds the modulated bilinear im2col tile
// cols[pixel][tap*Cin + ci] in shared memory; phase 2 is a small GEMM against w[Cout][9*Cin].
constexpr int DCN_PIX = 32;
constexpr int DCN_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(DCN_THREADS) dcn3x3_kernel(const T* __restrict__ x, int x_cs,
                                                             const T* __restrict__ off, int off_cs,
                                                             const T* __restrict__ msk, int msk_cs, int mask_is_logit,
                                                             const T* __restrict__ w, T* __restrict__ y, int y_cs,
                                                             int H, int W, int Cin, int Cout, long long npix) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float cols[];  // [DCN_PIX][K + 1]
    const int K = 9 * Cin;
    const int ld = K + 1;
    const long long p0 = (long long)blockIdx.x * DCN_PIX;
    // phase 1: thread <-> (pixel, tap, ci) with ci fastest
    for (int e = threadIdx.x; e < DCN_PIX * K; e += DCN_THREADS) {
        const int ci = e % Cin;
        const int tap = (e / Cin) % 9;
        const int pl = e / K;
        const long long pix = p0 + pl;
        float v = 0.f;
        if (pix < npix) {
            const int wq = (int)(pix % W);
            const int hq = (int)((pix / W) % H);
            const long long n = pix / ((long long)W * H);
            const T* ofp = off + pix * off_cs;
            const float dy = ldf(ofp + 2 * tap), dx = ldf(ofp + 2 * tap + 1);
            float m = ldf(msk + pix * msk_cs + tap);
            if (mask_is_logit) m = sigmoidf_(m);
            const float py = (float)(hq + tap / 3 - 1) + dy;
            const float px = (float)(wq + tap % 3 - 1) + dx;
            if (py > -1.f && py < (float)H && px > -1.f && px < (float)W) {
                const int y0 = (int)floorf(py), x0 = (int)floorf(px);
                const float ly = py - (float)y0, lx = px - (float)x0;
                const float hy = 1.f - ly, hx = 1.f - lx;
                const T* xn = x + n * (long long)H * W * x_cs + ci;
                float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
                if (y0 >= 0 && x0 >= 0) v00 = ldf(xn + ((long long)y0 * W + x0) * x_cs);
                if (y0 >= 0 && x0 + 1 <= W - 1) v01 = ldf(xn + ((long long)y0 * W + x0 + 1) * x_cs);
                if (y0 + 1 <= H - 1 && x0 >= 0) v10 = ldf(xn + ((long long)(y0 + 1) * W + x0) * x_cs);
                if (y0 + 1 <= H - 1 && x0 + 1 <= W - 1) v11 = ldf(xn + ((long long)(y0 + 1) * W + x0 + 1) * x_cs);
                v = (hy * hx * v00 + hy * lx * v01 + ly * hx * v10 + ly * lx * v11) * m;
            }
        }
        cols[pl * ld + tap * Cin + ci] = v;
    }
    __syncthreads();
    // phase 2: thread <-> (pixel, cout) pairs, cout fastest
    for (int e = threadIdx.x; e < DCN_PIX * Cout; e += DCN_THREADS) {
        const int co = e % Cout, pl = e / Cout;
        const long long pix = p0 + pl;
        if (pix >= npix) continue;
        const T* wr = w + (long long)co * K;
        const float* cr = cols + pl * ld;
        float a = 0.f;
        for (int k = 0; k < K; ++k) a = fmaf(cr[k], ldf(wr + k), a);
        stf(y + pix * y_cs + co, a);
    }
}

// ------------------------------------------------------------------ decode
struct DecodeLevels {
    const void* raw[4];
    int H[4], W[4], cs[4], a0[4];
    float stride[4];
    int nl, A;
};

template <typename T>
__global__ void __launch_bounds__(128) decode_kernel(DecodeLevels L, int reg_max, int nc, int dist_only,
                                                     float* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (a >= L.A) return;
    int l = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < L.nl && a >= L.a0[i]) l = i;
    const int la = a - L.a0[l];
    const int Wl = L.W[l], Hl = L.H[l];
    const int hq = la / Wl, wq = la - hq * Wl;
    const T* p = (const T*)L.raw[l] + ((long long)n * Hl * Wl + la) * L.cs[l];
    float d[4];
    for (int side = 0; side < 4; ++side) {
        if (reg_max > 1) {
            // softmax over reg_max bins, expectation with weights 0..reg_max-1 (DFL)
            float mx = -INFINITY;
            for (int k = 0; k < reg_max; ++k) mx = fmaxf(mx, ldf(p + side * reg_max + k));
            float den = 0.f, num = 0.f;
            for (int k = 0; k < reg_max; ++k) {
                const float e = expf(ldf(p + side * reg_max + k) - mx);
                den += e;
                num += e * (float)k;
            }
            d[side] = num / den;
        } else {
            d[side] = ldf(p + side);
        }
    }
    if (dist_only) {  // DFL.forward alone (block.py:50-53): (N, 4, A) expectations
        float* yd = y + (long long)n * 4 * L.A + a;
        for (int side = 0; side < 4; ++side) yd[(long long)side * L.A] = d[side];
        return;
    }
    const float ax = (float)wq + 0.5f, ay = (float)hq + 0.5f, st = L.stride[l];
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    float* yo = y + (long long)n * (4 + nc) * L.A + a;
    yo[0] = (x1 + x2) / 2.f * st;
    yo[(long long)L.A] = (y1 + y2) / 2.f * st;
    yo[2LL * L.A] = (x2 - x1) * st;
    yo[3LL * L.A] = (y2 - y1) * st;
    const T* pc = p + 4 * reg_max;
    for (int j = 0; j < nc; ++j) yo[(4LL + j) * L.A] = sigmoidf_(ldf(pc + j));
}

// bf16 variant that stages the block's 128 anchor rows through shared memory: the row-per-thread pattern of
// decode_kernel issues 2-byte loads at a 132-byte stride (66 channels); here the block copies its contiguous span
// with coalesced 4-byte loads and each thread then reads its own row from shared memory (odd word pitch, no bank
// conflicts).  Same arithmetic, same order of operations as decode_kernel.
__global__ void __launch_bounds__(128) decode_staged_kernel(DecodeLevels L, int reg_max, int nc, int dist_only,
                                                            float* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ uint32_t dec_sm[];
    const int a_blk = blockIdx.x * 128, n = blockIdx.y;
    const int no = 4 * reg_max + nc, nw = no / 2, pitch = nw | 1;
    int l = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < L.nl && a_blk >= L.a0[i]) l = i;
    const int Wl = L.W[l], Hl = L.H[l];
    const int la0 = a_blk - L.a0[l];
    const int cnt = min(128, Hl * Wl - la0);          // the host only takes this path when blocks do not straddle levels
    const __nv_bfloat16* base = (const __nv_bfloat16*)L.raw[l] + ((long long)n * Hl * Wl + la0) * L.cs[l];
    for (int e = threadIdx.x; e < cnt * nw; e += 128) {
        const int r = e / nw, c = e - r * nw;
        dec_sm[r * pitch + c] = __ldg(reinterpret_cast<const uint32_t*>(base + (long long)r * L.cs[l]) + c);
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= cnt) return;
    const int a = a_blk + t, la = la0 + t;
    const int hq = la / Wl, wq = la - hq * Wl;
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(dec_sm + t * pitch);
    float d[4];
    if (reg_max == 16) {
        // the common head (TOODHead / stock Detect at reg_max 16): fully unrolled, bins read as 32-bit pairs from the
        // staged row (the generic loop below costs ~2,400 instructions per anchor: ncu SM 50 % for 27 us).  Same
        // operations in the same order as the loop: running max, then exp / sum / weighted sum in bin order.
        const uint32_t* pw = dec_sm + t * pitch;
#pragma unroll
        for (int side = 0; side < 4; ++side) {
            float v[16];
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) {
                const uint32_t w = pw[side * 8 + k2];
                v[2 * k2] = __uint_as_float(w << 16);
                v[2 * k2 + 1] = __uint_as_float(w & 0xffff0000u);
            }
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 16; ++k) mx = fmaxf(mx, v[k]);
            float den = 0.f, num = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const float e = __expf(v[k] - mx);
                den += e;
                num += e * (float)k;
            }
            d[side] = num / den;
        }
    } else
    for (int side = 0; side < 4; ++side) {
        if (reg_max > 1) {
            float mx = -INFINITY;
            for (int k = 0; k < reg_max; ++k) mx = fmaxf(mx, __bfloat162float(p[side * reg_max + k]));
            float den = 0.f, num = 0.f;
            for (int k = 0; k < reg_max; ++k) {
                const float e = __expf(__bfloat162float(p[side * reg_max + k]) - mx);   // ex2.approx: 2^-21 relative, bf16 inputs
                den += e;
                num += e * (float)k;
            }
            d[side] = num / den;
        } else {
            d[side] = __bfloat162float(p[side]);
        }
    }
    if (dist_only) {
        float* yd = y + (long long)n * 4 * L.A + a;
        for (int side = 0; side < 4; ++side) yd[(long long)side * L.A] = d[side];
        return;
    }
    const float ax = (float)wq + 0.5f, ay = (float)hq + 0.5f, st = L.stride[l];
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    float* yo = y + (long long)n * (4 + nc) * L.A + a;
    yo[0] = (x1 + x2) / 2.f * st;
    yo[(long long)L.A] = (y1 + y2) / 2.f * st;
    yo[2LL * L.A] = (x2 - x1) * st;
    yo[3LL * L.A] = (y2 - y1) * st;
    const __nv_bfloat16* pc = p + 4 * reg_max;
    for (int j = 0; j < nc; ++j) yo[(4LL + j) * L.A] = sigmoidf_(__bfloat162float(pc[j]));
}

// reg_max = 16 heads with at most 8 classes and 16-byte aligned rows (the TOOD head of the full config: 66 channels at a
// stride of 72): every thread reads ITS anchor's row with nine 16-byte loads straight into registers -- no staging, no
// index division, no barrier -- and walks it fully unrolled.  Same operations in the same order as decode_kernel.
__global__ void __launch_bounds__(128) decode_rows16_kernel(DecodeLevels L, int nc, int dist_only, float* __restrict__ y) {
    pdl_trigger();
    pdl_wait();
    const int a = blockIdx.x * 128 + threadIdx.x, n = blockIdx.y;
    if (a >= L.A) return;
    int l = 0;
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (i < L.nl && a >= L.a0[i]) l = i;
    const int Wl = L.W[l], Hl = L.H[l], la = a - L.a0[l];
    const int hq = la / Wl, wq = la - hq * Wl;
    const uint4* row = reinterpret_cast<const uint4*>((const __nv_bfloat16*)L.raw[l] + ((long long)n * Hl * Wl + la) * L.cs[l]);
    uint32_t w[36];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        const uint4 v = __ldg(row + c);
        w[4 * c] = v.x; w[4 * c + 1] = v.y; w[4 * c + 2] = v.z; w[4 * c + 3] = v.w;
    }
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        float v[16];
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) {
            v[2 * k2] = __uint_as_float(w[side * 8 + k2] << 16);
            v[2 * k2 + 1] = __uint_as_float(w[side * 8 + k2] & 0xffff0000u);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) mx = fmaxf(mx, v[k]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float e = __expf(v[k] - mx);
            den += e;
            num += e * (float)k;
        }
        d[side] = num / den;
    }
    if (dist_only) {
        float* yd = y + (long long)n * 4 * L.A + a;
#pragma unroll
        for (int side = 0; side < 4; ++side) yd[(long long)side * L.A] = d[side];
        return;
    }
    const float ax = (float)wq + 0.5f, ay = (float)hq + 0.5f, st = L.stride[l];
    const float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3];
    float* yo = y + (long long)n * (4 + nc) * L.A + a;
    yo[0] = (x1 + x2) / 2.f * st;
    yo[(long long)L.A] = (y1 + y2) / 2.f * st;
    yo[2LL * L.A] = (x2 - x1) * st;
    yo[3LL * L.A] = (y2 - y1) * st;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nc) {
            const uint32_t ww = w[32 + (j >> 1)];
            yo[(4LL + j) * L.A] = sigmoidf_(__uint_as_float((j & 1) ? (ww & 0xffff0000u) : (ww << 16)));
        }
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_dwconv7_ln(const void* x, int x_cs, const void* w, const float* bias, const float* ln_w,
                               const float* ln_b, float eps, void* y, int y_cs, int N, int H, int W, int C, int dtype,
                               void* stream) {
    MGDT_CHECK(x && w && bias && ln_w && ln_b && y, "dwconv7_ln: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && C <= 32 * DW_MAXPL, "dwconv7_ln: C=%d unsupported (max %d)", C,
               32 * DW_MAXPL);
    MGDT_CHECK(x_cs >= C && y_cs >= C, "dwconv7_ln: bad strides");
    const long long npix = (long long)N * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    if (C <= 128) {  // shared-memory tiled kernel, CPL = ceil(C / 32) channels per lane
        const int cpl = (C + 31) / 32;
        int rc = 0;
        MGDT_DTYPE_SWITCH(dtype, T, {
            if (cpl == 1) rc = launch_dw_tiled<T, 1>(x, x_cs, w, bias, ln_w, ln_b, eps, y, y_cs, N, H, W, C, s);
            else if (cpl == 2) rc = launch_dw_tiled<T, 2>(x, x_cs, w, bias, ln_w, ln_b, eps, y, y_cs, N, H, W, C, s);
            else if (cpl == 3) rc = launch_dw_tiled<T, 3>(x, x_cs, w, bias, ln_w, ln_b, eps, y, y_cs, N, H, W, C, s);
            else rc = launch_dw_tiled<T, 4>(x, x_cs, w, bias, ln_w, ln_b, eps, y, y_cs, N, H, W, C, s);
        });
        if (rc) return rc;
        MGDT_LAUNCH_CHECK("dwconv7_ln_tiled");
        return 0;
    }
    MGDT_DTYPE_SWITCH(dtype, T, {
        launch_k(dwconv7_ln_kernel<T>, dim3(cdiv(npix, DW_WARPS)), dim3(DW_WARPS * 32), 0, s, (const T*)x, x_cs, (const T*)w, bias, ln_w, ln_b, eps, (T*)y, y_cs, H, W, C, npix);
    });
    MGDT_LAUNCH_CHECK("dwconv7_ln");
    return 0;
}

#ifdef MGDT_WITH_UMMA
namespace mgdt {
bool dcn_umma_supported(const void* x, int x_cs, const void* w_umma, int N, int H, int W, int Cin, int Cout);
int dcn_umma(const void* x, int x_cs, const void* offset, int off_cs, const void* mask, int mask_cs, int mask_is_logit,
             const void* w_umma, int w_f16, void* y, int y_cs, int N, int H, int W, int Cin, int Cout, void* stat_acc, int stat_q,
             int stat_sq, int stat_copies, cudaStream_t s);
}
#endif

extern "C" int mgdt_dcn3x3_path(const void* x, int x_cs, const void* w_umma, int N, int H, int W, int Cin, int Cout, int dtype) {
#ifdef MGDT_WITH_UMMA
    if (dtype == MGDT_BF16 && dcn_umma_supported(x, x_cs, w_umma, N, H, W, Cin, Cout)) return 2;
#endif
    return 1;
}

extern "C" int mgdt_dcn3x3(const void* x, int x_cs, const void* offset, int off_cs, const void* mask, int mask_cs,
                           int mask_is_logit, const void* w, const void* w_umma, int w_umma_f16, void* y, int y_cs, int N,
                           int H, int W, int Cin, int Cout, int dtype, void* stat_acc, int stat_q, int stat_sq, int stat_copies,
                           void* stream) {
    MGDT_CHECK(x && offset && mask && w && y, "dcn3x3: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && off_cs >= 18 && mask_cs >= 9, "dcn3x3: bad shape");
    MGDT_CHECK(!stat_acc || ((stat_q == 0 || stat_q == 1 || stat_q == 5) && stat_q + (stat_sq ? 1 : 0) > 0 && ((uintptr_t)stat_acc & 7) == 0),
               "dcn3x3: bad fused-statistics request");
#ifdef MGDT_WITH_UMMA
    if (dtype == MGDT_BF16 && dcn_umma_supported(x, x_cs, w_umma, N, H, W, Cin, Cout))
        return dcn_umma(x, x_cs, offset, off_cs, mask, mask_cs, mask_is_logit, w_umma, w_umma_f16, y, y_cs, N, H, W, Cin, Cout,
                        stat_acc, stat_q, stat_sq, stat_copies, (cudaStream_t)stream);
#endif
    if (stat_acc) return set_error(-ENOTSUP, "dcn3x3: fused statistics need the tcgen05 path (see mgdt_dcn3x3_path)");
    const size_t smem = sizeof(float) * DCN_PIX * (9 * Cin + 1);
    MGDT_CHECK(smem <= 200 * 1024, "dcn3x3: Cin=%d too large", Cin);
    const long long npix = (long long)N * H * W;
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_DTYPE_SWITCH(dtype, T, {
        if (smem + 1024 > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(dcn3x3_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return set_error(-EIO, "dcn3x3: smem attr: %s", cudaGetErrorString(e));
        }
        launch_k(dcn3x3_kernel<T>, dim3(cdiv(npix, DCN_PIX)), dim3(DCN_THREADS), smem, s, (const T*)x, x_cs, (const T*)offset, off_cs,
                                                                        (const T*)mask, mask_cs, mask_is_logit,
                                                                        (const T*)w, (T*)y, y_cs, H, W, Cin, Cout, npix);
    });
    MGDT_LAUNCH_CHECK("dcn3x3");
    return 0;
}

extern "C" int mgdt_decode(const mgdt_decode_level* levels, int nl, int N, int reg_max, int nc, int dist_only, float* y,
                           int dtype, void* stream) {
    MGDT_CHECK(levels && y, "decode: null pointer");
    MGDT_CHECK(nl >= 1 && nl <= 4, "decode: 1..4 levels supported, got %d", nl);
    MGDT_CHECK(N > 0 && reg_max >= 1 && reg_max <= 64 && nc >= 0, "decode: bad shape");
    DecodeLevels L;
    L.nl = nl;
    int a0 = 0;
    for (int i = 0; i < 4; ++i) {
        if (i < nl) {
            MGDT_CHECK(levels[i].raw && levels[i].H > 0 && levels[i].W > 0 && levels[i].cs >= 4 * reg_max + nc,
                       "decode: bad level %d", i);
            L.raw[i] = levels[i].raw; L.H[i] = levels[i].H; L.W[i] = levels[i].W; L.cs[i] = levels[i].cs;
            L.stride[i] = levels[i].stride; L.a0[i] = a0;
            a0 += levels[i].H * levels[i].W;
        } else {
            L.raw[i] = nullptr; L.H[i] = L.W[i] = L.cs[i] = 0; L.stride[i] = 0.f; L.a0[i] = 0x7fffffff;
        }
    }
    L.A = a0;
    // staged bf16 path: even channel count, 4-byte aligned rows, every level a multiple of 128 anchors (so no block
    // straddles two levels) or a single level
    bool staged = dtype == MGDT_BF16 && ((4 * reg_max + nc) & 1) == 0;
    for (int i = 0; i < nl && staged; ++i) {
        if ((levels[i].cs & 1) || ((uintptr_t)levels[i].raw & 3)) staged = false;
        if (i + 1 < nl && (levels[i].H * levels[i].W) % 128 != 0) staged = false;
    }
    // registers-only path (decode_rows16_kernel): reg_max 16, at most 8 classes, rows of >= 72 elements at 16-byte alignment
    bool rows16 = dtype == MGDT_BF16 && reg_max == 16 && nc <= 8;
    for (int i = 0; i < nl && rows16; ++i)
        if (levels[i].cs < 72 || (levels[i].cs & 7) || ((uintptr_t)levels[i].raw & 15)) rows16 = false;
    if (rows16) {
        launch_k(decode_rows16_kernel, dim3(cdiv(L.A, 128), N), dim3(128), 0, (cudaStream_t)stream, L, nc, dist_only, y);
        MGDT_LAUNCH_CHECK("decode_rows16");
        return 0;
    }
    if (staged) {
        const int nw = (4 * reg_max + nc) / 2;
        launch_k(decode_staged_kernel, dim3(cdiv(L.A, 128), N), dim3(128), (size_t)128 * (nw | 1) * 4, (cudaStream_t)stream, L,
                 reg_max, nc, dist_only, y);
        MGDT_LAUNCH_CHECK("decode_staged");
        return 0;
    }
    MGDT_DTYPE_SWITCH(dtype, T, {
        launch_k(decode_kernel<T>, dim3(dim3(cdiv(L.A, 128), N)), dim3(128), 0, (cudaStream_t)stream, L, reg_max, nc, dist_only, y);
    });
    MGDT_LAUNCH_CHECK("decode");
    return 0;
}
