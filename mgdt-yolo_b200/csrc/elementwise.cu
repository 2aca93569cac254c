// Memory-bound NHWC kernels: per-(n,c) affine + activation, resampling into concat slices, the
// SPPF pooling chain, the injection gate and input preprocessing.
//
// Mapping: one thread owns VEC = 8 consecutive channels of one pixel (a 16-byte bf16 / 32-byte fp32
// access), the channel-vector index is the fastest-varying thread index, so a warp touches
// contiguous bytes; all index arithmetic is 32-bit.  A scalar (VEC = 1) instantiation covers
// channel counts / strides / pointers that are not 8-aligned.
#include "common.cuh"

namespace mgdt {

constexpr int EW_THREADS = 256;

static inline int ew_grid(long long total, int waves = 4) {
    long long b = (total + EW_THREADS - 1) / EW_THREADS;
    const long long cap = 148LL * 8 * waves;   // 8 x 256 threads are resident per SM; kernels grid-stride beyond that
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ---- vector load/store of V channels as floats
template <typename T, int V> struct VecIO;
template <typename T> struct VecIO<T, 1> {
    static __device__ __forceinline__ void ld(const T* p, float* f) { f[0] = ldf(p); }
    static __device__ __forceinline__ void st(T* p, const float* f) { stf(p, f[0]); }
};
template <> struct VecIO<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};
template <> struct VecIO<float, 8> {
    static __device__ __forceinline__ void ld(const float* p, float* f) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void st(float* p, const float* f) {
        reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
};

__device__ __forceinline__ void unpack_bf8(const uint4& v, float* f) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
}

static inline bool aligned8(const void* p, int cs, size_t esize) {
    return p == nullptr || ((((uintptr_t)p) % (8 * esize)) == 0 && (cs % 8) == 0);
}

// ------------------------------------------------------------------ affine + act (+ other)
template <typename T, int V>
__global__ void __launch_bounds__(EW_THREADS) affine_act_kernel(const T* __restrict__ x, int x_cs,
                                                                const float* __restrict__ a,
                                                                const float* __restrict__ b,
                                                                const T* __restrict__ other, int o_cs, int act,
                                                                T* __restrict__ y, int y_cs, unsigned HW, unsigned C,
                                                                unsigned total) {
    pdl_trigger();
    pdl_wait();
    const unsigned CV = C / V;
    const unsigned stride = gridDim.x * EW_THREADS;
    // two independent elements per iteration (all loads first), per-(n,c) scales as vector loads, activation switch
    // outside the unrolled loops (SFU forms for bf16 storage, exact forms in the fp32 validation mode)
    for (unsigned i0 = blockIdx.x * EW_THREADS + threadIdx.x; i0 < total; i0 += 2 * stride) {
        float f[2][V], g[2][V], sa[2][V], sb[2][V];
        unsigned pixs[2], cs_[2];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const unsigned i = i0 + u * stride;
            on[u] = i < total;
            if (!on[u]) continue;
            const unsigned pix = i / CV, cv = i - pix * CV;
            const unsigned c = cv * V;
            pixs[u] = pix; cs_[u] = c;
            VecIO<T, V>::ld(x + (size_t)pix * x_cs + c, f[u]);
            if (other) VecIO<T, V>::ld(other + (size_t)pix * o_cs + c, g[u]);
            if (a || b) {
                const unsigned n = pix / HW;
                if (a) VecIO<float, V>::ld(a + (size_t)n * C + c, sa[u]);
                if (b) VecIO<float, V>::ld(b + (size_t)n * C + c, sb[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
#pragma unroll
            for (int j = 0; j < V; ++j) {
                if (a) f[u][j] *= sa[u][j];
                if (b) f[u][j] += sb[u][j];
            }
            if (sizeof(T) == 2) {
                switch (act) {
#define MGDT_ACT_CASE(A) case A: _Pragma("unroll") for (int j = 0; j < V; ++j) f[u][j] = act_fast<A>(f[u][j]); break;
                    MGDT_ACT_CASE(MGDT_ACT_SILU)
                    MGDT_ACT_CASE(MGDT_ACT_RELU)
                    MGDT_ACT_CASE(MGDT_ACT_SIGMOID)
                    MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
                    MGDT_ACT_CASE(MGDT_ACT_GELU)
#undef MGDT_ACT_CASE
                    default: break;
                }
            } else if (act != MGDT_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < V; ++j) f[u][j] = apply_act(f[u][j], act);
            }
            if (other) {
#pragma unroll
                for (int j = 0; j < V; ++j) f[u][j] += g[u][j];
            }
            VecIO<T, V>::st(y + (size_t)pixs[u] * y_cs + cs_[u], f[u]);
        }
    }
}

// bf16, 8-channel chunks: four chunks per thread are requested (packed, 16 bytes each, + the `other` operand) before any
// arithmetic, the per-(n,c) scales are fetched when a chunk is processed (L1 hits), <= 64 registers so that four
// CTAs share an SM: ~2x the bytes in flight of the generic kernel, which was latency-bound at ~45 % of the HBM roofline.
template <bool HAS_OTHER>
__global__ void __launch_bounds__(EW_THREADS, 4) affine_act_bf16x4_kernel(const __nv_bfloat16* __restrict__ x, int x_cs,
                                                                         const float* __restrict__ a, const float* __restrict__ b,
                                                                         const __nv_bfloat16* __restrict__ other, int o_cs, int act,
                                                                         __nv_bfloat16* __restrict__ y, int y_cs, unsigned HW,
                                                                         unsigned C, unsigned total) {
    pdl_trigger();
    pdl_wait();
    const unsigned CV = C / 8;
    const unsigned stride = gridDim.x * EW_THREADS;
    for (unsigned i0 = blockIdx.x * EW_THREADS + threadIdx.x; i0 < total; i0 += 4 * stride) {
        uint4 xv[4], ov[HAS_OTHER ? 4 : 1];
        unsigned pixs[4], cs_[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = min(i0 + u * stride, total - 1);
            const unsigned pix = i / CV, cv = i - pix * CV;
            pixs[u] = pix; cs_[u] = cv * 8;
            xv[u] = __ldg(reinterpret_cast<const uint4*>(x + (size_t)pix * x_cs + cv * 8));
            if (HAS_OTHER) ov[u] = __ldg(reinterpret_cast<const uint4*>(other + (size_t)pix * o_cs + cv * 8));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + u * stride >= total) break;
            float f[8];
            const uint32_t* w = reinterpret_cast<const uint32_t*>(&xv[u]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { f[2 * j] = __uint_as_float(w[j] << 16); f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u); }
            if (a || b) {
                const unsigned n = pixs[u] / HW;
                if (a) {
                    const float4* sp = reinterpret_cast<const float4*>(a + (size_t)n * C + cs_[u]);
                    const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                    f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w; f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
                }
                if (b) {
                    const float4* sp = reinterpret_cast<const float4*>(b + (size_t)n * C + cs_[u]);
                    const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                    f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w; f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
                }
            }
            switch (act) {
#define MGDT_ACT_CASE(A) case A: _Pragma("unroll") for (int j = 0; j < 8; ++j) f[j] = act_fast<A>(f[j]); break;
                MGDT_ACT_CASE(MGDT_ACT_SILU)
                MGDT_ACT_CASE(MGDT_ACT_RELU)
                MGDT_ACT_CASE(MGDT_ACT_SIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_HSIGMOID)
                MGDT_ACT_CASE(MGDT_ACT_GELU)
#undef MGDT_ACT_CASE
                default: break;
            }
            if (HAS_OTHER) {
                const uint32_t* ow = reinterpret_cast<const uint32_t*>(&ov[HAS_OTHER ? u : 0]);
#pragma unroll
                for (int j = 0; j < 4; ++j) { f[2 * j] += __uint_as_float(ow[j] << 16); f[2 * j + 1] += __uint_as_float(ow[j] & 0xffff0000u); }
            }
            VecIO<__nv_bfloat16, 8>::st(y + (size_t)pixs[u] * y_cs + cs_[u], f);
        }
    }
}

// ------------------------------------------------------------------ resample
__device__ __forceinline__ void bilinear_src(int o, int in, int out, int& i0, int& i1, float& l) {
    // F.interpolate(mode='bilinear', align_corners=False): src = (o + 0.5) * in/out - 0.5, clamped at 0
    const float scale = (float)in / (float)out;
    float src = ((float)o + 0.5f) * scale - 0.5f;
    if (src < 0.f) src = 0.f;
    i0 = (int)src;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l = src - (float)i0;
}

template <typename T, int V>
__global__ void __launch_bounds__(EW_THREADS) resample_kernel(const T* __restrict__ x, int x_cs, int Hi, int Wi,
                                                              T* __restrict__ y, int y_cs, int Ho, int Wo, unsigned C,
                                                              int mode, unsigned total) {
    pdl_trigger();
    pdl_wait();
    const unsigned CV = C / V;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, pix = i / CV;
        const unsigned c = cv * V;
        const int wo = (int)(pix % (unsigned)Wo);
        const unsigned t = pix / (unsigned)Wo;
        const int ho = (int)(t % (unsigned)Ho);
        const unsigned n = t / (unsigned)Ho;
        const T* xn = x + (size_t)n * Hi * Wi * x_cs + c;
        float f[V];
        if (mode == MGDT_RS_COPY) {
            VecIO<T, V>::ld(xn + (size_t)(ho * Wi + wo) * x_cs, f);
        } else if (mode == MGDT_RS_NEAREST) {
            // nn.Upsample(mode='nearest'): src = floor(dst * in/out)
            const int hi = min((int)floorf((float)ho * ((float)Hi / (float)Ho)), Hi - 1);
            const int wi = min((int)floorf((float)wo * ((float)Wi / (float)Wo)), Wi - 1);
            VecIO<T, V>::ld(xn + (size_t)(hi * Wi + wi) * x_cs, f);
        } else if (mode == MGDT_RS_AVGPOOL) {
            // adaptive_avg_pool2d: window [floor(o*in/out), ceil((o+1)*in/out))
            const int h0 = (ho * Hi) / Ho, h1 = ((ho + 1) * Hi + Ho - 1) / Ho;
            const int w0 = (wo * Wi) / Wo, w1 = ((wo + 1) * Wi + Wo - 1) / Wo;
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = 0.f;
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) {
                    float g[V];
                    VecIO<T, V>::ld(xn + (size_t)(h * Wi + w) * x_cs, g);
#pragma unroll
                    for (int j = 0; j < V; ++j) f[j] += g[j];
                }
            const float cnt = (float)((h1 - h0) * (w1 - w0));
#pragma unroll
            for (int j = 0; j < V; ++j) f[j] = f[j] / cnt;
        } else {
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(ho, Hi, Ho, h0, h1, lh);
            bilinear_src(wo, Wi, Wo, w0, w1, lw);
            float v00[V], v01[V], v10[V], v11[V];
            VecIO<T, V>::ld(xn + (size_t)(h0 * Wi + w0) * x_cs, v00);
            VecIO<T, V>::ld(xn + (size_t)(h0 * Wi + w1) * x_cs, v01);
            VecIO<T, V>::ld(xn + (size_t)(h1 * Wi + w0) * x_cs, v10);
            VecIO<T, V>::ld(xn + (size_t)(h1 * Wi + w1) * x_cs, v11);
#pragma unroll
            for (int j = 0; j < V; ++j)
                f[j] = (1.f - lh) * ((1.f - lw) * v00[j] + lw * v01[j]) + lh * ((1.f - lw) * v10[j] + lw * v11[j]);
        }
        VecIO<T, V>::st(y + (size_t)pix * y_cs + c, f);
    }
}

// bf16 fast paths of the two resampling shapes the GD neck uses (same arithmetic and order as resample_kernel, so the
// results are identical): exact K x K average pooling (Hi == K*Ho, Wi == K*Wo, K = 2 or 4) with all loads of a window
// requested before the first add, and exact 2x bilinear upsampling with one thread per 2x2 output block (the four
// source corners are shared by the block: 1 load per output instead of 4), as inject2x does.
template <int K>
__global__ void __launch_bounds__(EW_THREADS, 4) avgpool_exact_bf16_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int Wi,
                                                                          __nv_bfloat16* __restrict__ y, int y_cs, int Ho, int Wo,
                                                                          unsigned C, unsigned total) {
    pdl_trigger();
    pdl_wait();
    const unsigned CV = C / 8;
    const int Hi = K * Ho;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, pix = i / CV;
        const int wo = (int)(pix % (unsigned)Wo);
        const unsigned t = pix / (unsigned)Wo;
        const int ho = (int)(t % (unsigned)Ho);
        const unsigned n = t / (unsigned)Ho;
        const __nv_bfloat16* xw = x + ((size_t)n * Hi * Wi + (size_t)(ho * K) * Wi + wo * K) * x_cs + cv * 8;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
#pragma unroll
        for (int h0 = 0; h0 < K; h0 += 2) {   // two window rows (2K chunks) in flight at a time
            uint4 v[2][K];
#pragma unroll
            for (int dh = 0; dh < 2; ++dh)
#pragma unroll
                for (int w = 0; w < K; ++w) v[dh][w] = __ldg(reinterpret_cast<const uint4*>(xw + (size_t)((h0 + dh) * Wi + w) * x_cs));
#pragma unroll
            for (int dh = 0; dh < 2; ++dh)
#pragma unroll
                for (int w = 0; w < K; ++w) {
                    float g[8];
                    unpack_bf8(v[dh][w], g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] += g[j];
                }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] / (float)(K * K);
        VecIO<__nv_bfloat16, 8>::st(y + (size_t)pix * y_cs + cv * 8, f);
    }
}

__global__ void __launch_bounds__(EW_THREADS, 4) bilinear2x_bf16_kernel(const __nv_bfloat16* __restrict__ x, int x_cs, int Hi, int Wi,
                                                                       __nv_bfloat16* __restrict__ y, int y_cs, unsigned C,
                                                                       unsigned total) {
    pdl_trigger();
    pdl_wait();
    const int Ho = 2 * Hi, Wo = 2 * Wi;
    const unsigned CV = C / 8;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, blk = i / CV;
        const int bj = (int)(blk % (unsigned)(Wi + 1));
        const unsigned t = blk / (unsigned)(Wi + 1);
        const int bi = (int)(t % (unsigned)(Hi + 1));
        const unsigned n = t / (unsigned)(Hi + 1);
        const int ra = max(bi - 1, 0), rb = min(bi, Hi - 1), ca = max(bj - 1, 0), cb = min(bj, Wi - 1);
        const __nv_bfloat16* xn = x + (size_t)n * Hi * Wi * x_cs + cv * 8;
        const uint4 q00 = __ldg(reinterpret_cast<const uint4*>(xn + (size_t)(ra * Wi + ca) * x_cs));
        const uint4 q01 = __ldg(reinterpret_cast<const uint4*>(xn + (size_t)(ra * Wi + cb) * x_cs));
        const uint4 q10 = __ldg(reinterpret_cast<const uint4*>(xn + (size_t)(rb * Wi + ca) * x_cs));
        const uint4 q11 = __ldg(reinterpret_cast<const uint4*>(xn + (size_t)(rb * Wi + cb) * x_cs));
        float v00[8], v01[8], v10[8], v11[8];
        unpack_bf8(q00, v00); unpack_bf8(q01, v01); unpack_bf8(q10, v10); unpack_bf8(q11, v11);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ho = 2 * bi - 1 + (k >> 1), wo = 2 * bj - 1 + (k & 1);
            if (ho < 0 || ho >= Ho || wo < 0 || wo >= Wo) continue;
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(ho, Hi, Ho, h0, h1, lh);   // (h0, h1) == (ra, rb) whenever lh != 0
            bilinear_src(wo, Wi, Wo, w0, w1, lw);
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                f[j] = (1.f - lh) * ((1.f - lw) * v00[j] + lw * v01[j]) + lh * ((1.f - lw) * v10[j] + lw * v11[j]);
            VecIO<__nv_bfloat16, 8>::st(y + (((size_t)n * Ho + ho) * Wo + wo) * y_cs + cv * 8, f);
        }
    }
}

// ------------------------------------------------------------------ SPPF pooling chain
// maxpool(k,1,k/2) applied 1x/2x/3x == max over (k-1)*j+1 windows, j = 1..3 (padding is -inf).
template <typename T, int V>
__global__ void __launch_bounds__(EW_THREADS) sppf_pool_kernel(const T* __restrict__ x, int x_cs, T* __restrict__ y1,
                                                               T* __restrict__ y2, T* __restrict__ y3, int y_cs, int H,
                                                               int W, unsigned C, int r, unsigned total) {
    pdl_trigger();
    pdl_wait();
    const unsigned CV = C / V;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, pix = i / CV;
        const unsigned c = cv * V;
        const int w = (int)(pix % (unsigned)W);
        const unsigned t = pix / (unsigned)W;
        const int h = (int)(t % (unsigned)H);
        const unsigned n = t / (unsigned)H;
        const T* xn = x + (size_t)n * H * W * x_cs + c;
        float m1[V], m2[V], m3[V];
#pragma unroll
        for (int j = 0; j < V; ++j) m1[j] = m2[j] = m3[j] = -INFINITY;
        for (int dy = -3 * r; dy <= 3 * r; ++dy) {
            const int hh = h + dy;
            if (hh < 0 || hh >= H) continue;
            const int ady = dy < 0 ? -dy : dy;
            for (int dx = -3 * r; dx <= 3 * r; ++dx) {
                const int ww = w + dx;
                if (ww < 0 || ww >= W) continue;
                const int adx = dx < 0 ? -dx : dx;
                const int d = ady > adx ? ady : adx;
                float g[V];
                VecIO<T, V>::ld(xn + (size_t)(hh * W + ww) * x_cs, g);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    m3[j] = fmaxf(m3[j], g[j]);
                    if (d <= 2 * r) m2[j] = fmaxf(m2[j], g[j]);
                    if (d <= r) m1[j] = fmaxf(m1[j], g[j]);
                }
            }
        }
        VecIO<T, V>::st(y1 + (size_t)pix * y_cs + c, m1);
        VecIO<T, V>::st(y2 + (size_t)pix * y_cs + c, m2);
        VecIO<T, V>::st(y3 + (size_t)pix * y_cs + c, m3);
    }
}

// SPPF pooling chain for maps that fit in shared memory (bf16, 8-aligned): one CTA per (image, 8-channel chunk).
// pool_{k}(pool_{k}(x)) = pool_{2k-1}(x) with -inf padding, so the three outputs are three cascaded separable
// (row max, then column max) passes over a ping-pong pair of [H*W] x 16-byte tiles; the max of packed bf16 pairs
// is a selection, hence exact.
__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
    uint4 r;
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __hmax2(x[j], y[j]);
    return r;
}

__global__ void __launch_bounds__(EW_THREADS) sppf_pool_tile(const __nv_bfloat16* __restrict__ x, int x_cs,
                                                             __nv_bfloat16* __restrict__ y1, __nv_bfloat16* __restrict__ y2,
                                                             __nv_bfloat16* __restrict__ y3, int y_cs, int H, int W, int C8,
                                                             int r) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ uint4 sp_buf[];
    const int HW = H * W;
    uint4* A = sp_buf;
    uint4* B = sp_buf + HW;
    const int n = blockIdx.x / C8, c = (blockIdx.x - n * C8) * 8;
    const size_t pix0 = (size_t)n * HW;
    for (int i = threadIdx.x; i < HW; i += EW_THREADS) A[i] = __ldg(reinterpret_cast<const uint4*>(x + (pix0 + i) * x_cs + c));
    __syncthreads();
    __nv_bfloat16* outs[3] = {y1, y2, y3};
    for (int st = 0; st < 3; ++st) {
        for (int i = threadIdx.x; i < HW; i += EW_THREADS) {
            const int h = i / W, w = i - h * W;
            const int lo = max(w - r, 0), hi = min(w + r, W - 1);
            uint4 m = A[h * W + lo];
            for (int ww = lo + 1; ww <= hi; ++ww) m = max8(m, A[h * W + ww]);
            B[i] = m;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < HW; i += EW_THREADS) {
            const int h = i / W, w = i - h * W;
            const int lo = max(h - r, 0), hi = min(h + r, H - 1);
            uint4 m = B[lo * W + w];
            for (int hh = lo + 1; hh <= hi; ++hh) m = max8(m, B[hh * W + w]);
            A[i] = m;
            *reinterpret_cast<uint4*>(outs[st] + (pix0 + i) * y_cs + c) = m;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ injection gate
// relu6(x + 3) / 6: exact IEEE form in the fp32 validation mode, saturate(x/6 + 0.5) (2 instructions) for bf16
template <typename T> __device__ __forceinline__ float hsig(float v);
template <> __device__ __forceinline__ float hsig<float>(float v) { return fminf(fmaxf(v + 3.0f, 0.0f), 6.0f) / 6.0f; }
template <> __device__ __forceinline__ float hsig<__nv_bfloat16>(float v) { return __saturatef(fmaf(v, 1.0f / 6.0f, 0.5f)); }

template <typename T, int V>
__global__ void __launch_bounds__(EW_THREADS) inject_kernel(const T* __restrict__ local, int l_cs,
                                                            const T* __restrict__ gact, int a_cs,
                                                            const T* __restrict__ gfeat, int f_cs, T* __restrict__ y,
                                                            int y_cs, int H, int W, int Hg, int Wg, unsigned C,
                                                            unsigned total) {
    pdl_trigger();
    pdl_wait();
    const bool pool = H < Hg;
    const unsigned CV = C / V;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, pix = i / CV;
        const unsigned c = cv * V;
        const int w = (int)(pix % (unsigned)W);
        const unsigned t = pix / (unsigned)W;
        const int h = (int)(t % (unsigned)H);
        const unsigned n = t / (unsigned)H;
        const T* an = gact + (size_t)n * Hg * Wg * a_cs + c;
        const T* fn = gfeat + (size_t)n * Hg * Wg * f_cs + c;
        float sig[V], gf[V];
        if (pool) {
            const int h0 = (h * Hg) / H, h1 = ((h + 1) * Hg + H - 1) / H;
            const int w0 = (w * Wg) / W, w1 = ((w + 1) * Wg + W - 1) / W;
#pragma unroll
            for (int j = 0; j < V; ++j) sig[j] = gf[j] = 0.f;
            for (int hh = h0; hh < h1; ++hh)
                for (int ww = w0; ww < w1; ++ww) {
                    float a[V], f[V];
                    VecIO<T, V>::ld(an + (size_t)(hh * Wg + ww) * a_cs, a);
                    VecIO<T, V>::ld(fn + (size_t)(hh * Wg + ww) * f_cs, f);
#pragma unroll
                    for (int j = 0; j < V; ++j) { sig[j] += a[j]; gf[j] += f[j]; }
                }
            const float cnt = (float)((h1 - h0) * (w1 - w0));
#pragma unroll
            for (int j = 0; j < V; ++j) { sig[j] = sig[j] / cnt; gf[j] = gf[j] / cnt; }
        } else {
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(h, Hg, H, h0, h1, lh);
            bilinear_src(w, Wg, W, w0, w1, lw);
            const size_t o00 = (size_t)(h0 * Wg + w0), o01 = (size_t)(h0 * Wg + w1);
            const size_t o10 = (size_t)(h1 * Wg + w0), o11 = (size_t)(h1 * Wg + w1);
            float a00[V], a01[V], a10[V], a11[V], f00[V], f01[V], f10[V], f11[V];
            VecIO<T, V>::ld(an + o00 * a_cs, a00); VecIO<T, V>::ld(an + o01 * a_cs, a01);
            VecIO<T, V>::ld(an + o10 * a_cs, a10); VecIO<T, V>::ld(an + o11 * a_cs, a11);
            VecIO<T, V>::ld(fn + o00 * f_cs, f00); VecIO<T, V>::ld(fn + o01 * f_cs, f01);
            VecIO<T, V>::ld(fn + o10 * f_cs, f10); VecIO<T, V>::ld(fn + o11 * f_cs, f11);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                sig[j] = (1.f - lh) * ((1.f - lw) * hsig<T>(a00[j]) + lw * hsig<T>(a01[j])) +
                         lh * ((1.f - lw) * hsig<T>(a10[j]) + lw * hsig<T>(a11[j]));
                gf[j] = (1.f - lh) * ((1.f - lw) * f00[j] + lw * f01[j]) + lh * ((1.f - lw) * f10[j] + lw * f11[j]);
            }
        }
        float l[V];
        VecIO<T, V>::ld(local + (size_t)pix * l_cs + c, l);
#pragma unroll
        for (int j = 0; j < V; ++j) l[j] = l[j] * sig[j] + gf[j];
        VecIO<T, V>::st(y + (size_t)pix * y_cs + c, l);
    }
}

// ------------------------------------------------------------------ preprocess
template <typename S, typename T>
__global__ void __launch_bounds__(EW_THREADS) preprocess_kernel(const S* __restrict__ src, T* __restrict__ y, int y_cs,
                                                                int C, unsigned HW, float div, unsigned npix) {
    pdl_trigger();
    pdl_wait();
    // thread <-> pixel: per-plane reads are contiguous across the warp; the C outputs of a pixel are
    // adjacent, so the warp's stores cover one contiguous span.
    for (unsigned pix = blockIdx.x * EW_THREADS + threadIdx.x; pix < npix; pix += gridDim.x * EW_THREADS) {
        const unsigned n = pix / HW, p = pix - n * HW;
        const S* s = src + (size_t)n * C * HW + p;
        T* o = y + (size_t)pix * y_cs;
        for (int c = 0; c < C; ++c) stf(o + c, (float)s[(size_t)c * HW] / div);  // `im /= 255` (predictor.py:129)
    }
}

}  // namespace mgdt

using namespace mgdt;

#define MGDT_VEC_SWITCH(vec_ok, V, ...) \
    do {                                \
        if (vec_ok) {                   \
            constexpr int V = 8;        \
            __VA_ARGS__;                \
        } else {                        \
            constexpr int V = 1;        \
            __VA_ARGS__;                \
        }                               \
    } while (0)

extern "C" int mgdt_affine_act(const void* x, int x_cs, const float* a, const float* b, const void* other, int o_cs,
                               int act, void* y, int y_cs, int N, int H, int W, int C, int dtype, void* stream) {
    MGDT_CHECK(x && y, "affine_act: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && x_cs >= C && y_cs >= C, "affine_act: bad shape");
    MGDT_CHECK((long long)N * H * W * C < (1LL << 31), "affine_act: tensor too large for 32-bit indexing");
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && aligned8(x, x_cs, sizeof(T)) && aligned8(y, y_cs, sizeof(T)) &&
                         aligned8(other, other ? o_cs : 0, sizeof(T));
        MGDT_VEC_SWITCH(vec, V, {
            const unsigned total = (unsigned)((long long)N * H * W * (C / V));
            if constexpr (sizeof(T) == 2 && V == 8) {
                if (aligned8(a, 8, 4) && aligned8(b, 8, 4)) {
                    if (other) launch_k(affine_act_bf16x4_kernel<true>, dim3(ew_grid((total + 3) / 4, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, a, b, (const T*)other, o_cs, act, (T*)y, y_cs, (unsigned)(H * W), (unsigned)C, total);
                    else launch_k(affine_act_bf16x4_kernel<false>, dim3(ew_grid((total + 3) / 4, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, a, b, (const T*)other, o_cs, act, (T*)y, y_cs, (unsigned)(H * W), (unsigned)C, total);
                    MGDT_LAUNCH_CHECK("affine_act");
                    return 0;
                }
            }
            launch_k(affine_act_kernel<T, V>, dim3(ew_grid((total + 1) / 2, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, a, b, (const T*)other, o_cs, act, (T*)y, y_cs, (unsigned)(H * W), (unsigned)C, total);
        });
    });
    MGDT_LAUNCH_CHECK("affine_act");
    return 0;
}

extern "C" int mgdt_resample(const void* x, int x_cs, int Hi, int Wi, void* y, int y_cs, int Ho, int Wo, int N, int C,
                             int mode, int dtype, void* stream) {
    MGDT_CHECK(x && y, "resample: null pointer");
    MGDT_CHECK(N > 0 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && x_cs >= C && y_cs >= C, "resample: bad shape");
    MGDT_CHECK(mode >= 0 && mode <= 3, "resample: bad mode %d", mode);
    MGDT_CHECK(mode != MGDT_RS_COPY || (Hi == Ho && Wi == Wo), "resample: copy needs equal sizes");
    MGDT_CHECK((long long)N * Ho * Wo * C < (1LL << 31) && (long long)N * Hi * Wi * C < (1LL << 31),
               "resample: tensor too large for 32-bit indexing");
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && aligned8(x, x_cs, sizeof(T)) && aligned8(y, y_cs, sizeof(T));
        MGDT_VEC_SWITCH(vec, V, {
            if constexpr (sizeof(T) == 2 && V == 8) {
                const bool pool2 = mode == MGDT_RS_AVGPOOL && Hi == 2 * Ho && Wi == 2 * Wo;
                const bool pool4 = mode == MGDT_RS_AVGPOOL && Hi == 4 * Ho && Wi == 4 * Wo;
                if (pool2 || pool4) {
                    const unsigned tot = (unsigned)((long long)N * Ho * Wo * (C / 8));
                    if (pool2) launch_k(avgpool_exact_bf16_kernel<2>, dim3(ew_grid(tot, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, Wi, (T*)y, y_cs, Ho, Wo, (unsigned)C, tot);
                    else launch_k(avgpool_exact_bf16_kernel<4>, dim3(ew_grid(tot, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, Wi, (T*)y, y_cs, Ho, Wo, (unsigned)C, tot);
                    MGDT_LAUNCH_CHECK("resample");
                    return 0;
                }
                if (mode == MGDT_RS_BILINEAR && Ho == 2 * Hi && Wo == 2 * Wi && Hi > 1 && Wi > 1) {
                    const unsigned tot = (unsigned)((long long)N * (Hi + 1) * (Wi + 1) * (C / 8));
                    launch_k(bilinear2x_bf16_kernel, dim3(ew_grid(tot, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, Hi, Wi, (T*)y, y_cs, (unsigned)C, tot);
                    MGDT_LAUNCH_CHECK("resample");
                    return 0;
                }
            }
            const unsigned total = (unsigned)((long long)N * Ho * Wo * (C / V));
            launch_k(resample_kernel<T, V>, dim3(ew_grid(total)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, Hi, Wi, (T*)y, y_cs, Ho, Wo, (unsigned)C, mode, total);
        });
    });
    MGDT_LAUNCH_CHECK("resample");
    return 0;
}

extern "C" int mgdt_sppf_pool(const void* x, int x_cs, void* y1, void* y2, void* y3, int y_cs, int N, int H, int W,
                              int C, int k, int dtype, void* stream) {
    MGDT_CHECK(x && y1 && y2 && y3, "sppf_pool: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && (k & 1) && k >= 1, "sppf_pool: bad shape/k");
    MGDT_CHECK((long long)N * H * W * C < (1LL << 31), "sppf_pool: tensor too large for 32-bit indexing");
    if (dtype == MGDT_BF16 && C % 8 == 0 && aligned8(x, x_cs, 2) && aligned8(y1, y_cs, 2) && aligned8(y2, y_cs, 2) &&
        aligned8(y3, y_cs, 2) && (size_t)H * W * 32 <= 96 * 1024) {
        const size_t smem = (size_t)H * W * 32;
        cudaError_t e = cudaFuncSetAttribute(sppf_pool_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-EIO, "sppf_pool: smem attr: %s", cudaGetErrorString(e));
        launch_k(sppf_pool_tile, dim3(N * (C / 8)), dim3(EW_THREADS), smem, (cudaStream_t)stream, (const __nv_bfloat16*)x, x_cs, (__nv_bfloat16*)y1, (__nv_bfloat16*)y2, (__nv_bfloat16*)y3, y_cs, H, W, C / 8, k / 2);
        MGDT_LAUNCH_CHECK("sppf_pool_tile");
        return 0;
    }
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && aligned8(x, x_cs, sizeof(T)) && aligned8(y1, y_cs, sizeof(T)) &&
                         aligned8(y2, y_cs, sizeof(T)) && aligned8(y3, y_cs, sizeof(T));
        MGDT_VEC_SWITCH(vec, V, {
            const unsigned total = (unsigned)((long long)N * H * W * (C / V));
            launch_k(sppf_pool_kernel<T, V>, dim3(ew_grid(total)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)x, x_cs, (T*)y1, (T*)y2, (T*)y3, y_cs, H, W, (unsigned)C, k / 2, total);
        });
    });
    MGDT_LAUNCH_CHECK("sppf_pool");
    return 0;
}

namespace mgdt {
// Exact 2x bilinear upsampling (H == 2*Hg, W == 2*Wg, the GD neck's case): output rows {2*bi - 1, 2*bi} and columns
// {2*bj - 1, 2*bj} all interpolate between source rows {bi - 1, bi} and columns {bj - 1, bj} (clamped), so one thread
// loads the four corners of both global maps once (8 loads) and produces the whole 2x2 output block: 4 memory
// instructions per output chunk instead of 10.  Weights come from the same bilinear_src() as the general kernel, so
// the results are identical to it.
template <typename T, int V>
__global__ void __launch_bounds__(EW_THREADS) inject2x_kernel(const T* __restrict__ local, int l_cs, const T* __restrict__ gact,
                                                              int a_cs, const T* __restrict__ gfeat, int f_cs, T* __restrict__ y,
                                                              int y_cs, int Hg, int Wg, unsigned C, unsigned total) {
    pdl_trigger();
    pdl_wait();
    const int H = 2 * Hg, W = 2 * Wg;
    const unsigned CV = C / V;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, blk = i / CV;
        const unsigned c = cv * V;
        const int bj = (int)(blk % (unsigned)(Wg + 1));
        const unsigned t = blk / (unsigned)(Wg + 1);
        const int bi = (int)(t % (unsigned)(Hg + 1));
        const unsigned n = t / (unsigned)(Hg + 1);
        const int ra = max(bi - 1, 0), rb = min(bi, Hg - 1), ca = max(bj - 1, 0), cb = min(bj, Wg - 1);
        const T* an = gact + (size_t)n * Hg * Wg * a_cs + c;
        const T* fn = gfeat + (size_t)n * Hg * Wg * f_cs + c;
        float a[4][V], f[4][V];   // corners (ra,ca) (ra,cb) (rb,ca) (rb,cb)
        VecIO<T, V>::ld(an + (size_t)(ra * Wg + ca) * a_cs, a[0]); VecIO<T, V>::ld(an + (size_t)(ra * Wg + cb) * a_cs, a[1]);
        VecIO<T, V>::ld(an + (size_t)(rb * Wg + ca) * a_cs, a[2]); VecIO<T, V>::ld(an + (size_t)(rb * Wg + cb) * a_cs, a[3]);
        VecIO<T, V>::ld(fn + (size_t)(ra * Wg + ca) * f_cs, f[0]); VecIO<T, V>::ld(fn + (size_t)(ra * Wg + cb) * f_cs, f[1]);
        VecIO<T, V>::ld(fn + (size_t)(rb * Wg + ca) * f_cs, f[2]); VecIO<T, V>::ld(fn + (size_t)(rb * Wg + cb) * f_cs, f[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < V; ++j) a[k][j] = hsig<T>(a[k][j]);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int h = 2 * bi - 1 + dy;
            if (h < 0 || h >= H) continue;
            int h0, h1;
            float lh;
            bilinear_src(h, Hg, H, h0, h1, lh);   // (h0, h1) == (ra, rb) whenever lh != 0
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int w = 2 * bj - 1 + dx;
                if (w < 0 || w >= W) continue;
                int w0, w1;
                float lw;
                bilinear_src(w, Wg, W, w0, w1, lw);
                const size_t pix = ((size_t)n * H + h) * W + w;
                float l[V];
                VecIO<T, V>::ld(local + pix * l_cs + c, l);
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    const float sig = (1.f - lh) * ((1.f - lw) * a[0][j] + lw * a[1][j]) + lh * ((1.f - lw) * a[2][j] + lw * a[3][j]);
                    const float gf = (1.f - lh) * ((1.f - lw) * f[0][j] + lw * f[1][j]) + lh * ((1.f - lw) * f[2][j] + lw * f[3][j]);
                    l[j] = l[j] * sig + gf;
                }
                VecIO<T, V>::st(y + pix * y_cs + c, l);
            }
        }
    }
}
// bf16 form of inject2x_kernel: all twelve 16-byte loads of a 2x2 output block (8 corners + 4 local chunks) are issued
// before any arithmetic and kept packed, so a thread has 192 bytes in flight and three CTAs fit on an SM.  The
// arithmetic runs on channel PAIRS with packed fp32x2 instructions (the kernel was ALU-bound: ~40 scalar instructions
// per output element, 81 us for 228 MB): the four bilinear weights of an output pixel are shared by its channels, so
// sig and gfeat are one FMUL2 + three FFMA2 each.  pre_hsig: gact already holds h_sigmoid(global_act) -- applied in the
// producing conv's epilogue, BEFORE the interpolation exactly as the reference does (block.py:393).
__device__ __forceinline__ float2 bf2_unpack(uint32_t w32) { return make_float2(__uint_as_float(w32 << 16), __uint_as_float(w32 & 0xffff0000u)); }
__device__ __forceinline__ float2 hsig2(float2 v) {
    const float2 t = __ffma2_rn(v, make_float2(1.0f / 6.0f, 1.0f / 6.0f), make_float2(0.5f, 0.5f));
    return make_float2(__saturatef(t.x), __saturatef(t.y));
}
__global__ void __launch_bounds__(EW_THREADS, 3) inject2x_bf16_kernel(const __nv_bfloat16* __restrict__ local, int l_cs,
                                                                     const __nv_bfloat16* __restrict__ gact, int a_cs,
                                                                     const __nv_bfloat16* __restrict__ gfeat, int f_cs,
                                                                     __nv_bfloat16* __restrict__ y, int y_cs, int Hg, int Wg,
                                                                     unsigned C, unsigned total, int pre_hsig) {
    pdl_trigger();
    pdl_wait();
    using T = __nv_bfloat16;
    const int H = 2 * Hg, W = 2 * Wg;
    const unsigned CV = C / 8;
    for (unsigned i = blockIdx.x * EW_THREADS + threadIdx.x; i < total; i += gridDim.x * EW_THREADS) {
        const unsigned cv = i % CV, blk = i / CV;
        const unsigned c = cv * 8;
        const int bj = (int)(blk % (unsigned)(Wg + 1));
        const unsigned t = blk / (unsigned)(Wg + 1);
        const int bi = (int)(t % (unsigned)(Hg + 1));
        const unsigned n = t / (unsigned)(Hg + 1);
        const int ra = max(bi - 1, 0), rb = min(bi, Hg - 1), ca = max(bj - 1, 0), cb = min(bj, Wg - 1);
        const T* an = gact + (size_t)n * Hg * Wg * a_cs + c;
        const T* fn = gfeat + (size_t)n * Hg * Wg * f_cs + c;
        const int off[4] = {ra * Wg + ca, ra * Wg + cb, rb * Wg + ca, rb * Wg + cb};
        uint4 qa[4], qf[4], ql[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            qa[k] = __ldg(reinterpret_cast<const uint4*>(an + (size_t)off[k] * a_cs));
            qf[k] = __ldg(reinterpret_cast<const uint4*>(fn + (size_t)off[k] * f_cs));
        }
        size_t pix[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = 2 * bi - 1 + (k >> 1), w = 2 * bj - 1 + (k & 1);
            ok[k] = h >= 0 && h < H && w >= 0 && w < W;
            pix[k] = ((size_t)n * H + min(max(h, 0), H - 1)) * W + min(max(w, 0), W - 1);
            ql[k] = __ldg(reinterpret_cast<const uint4*>(local + pix[k] * l_cs + c));
        }
        // the gate's corner values as fp32 pairs, h-sigmoided once per corner (not once per output pixel)
        float2 ga[4][4], gf[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 av = bf2_unpack(reinterpret_cast<const uint32_t*>(&qa[q])[j]);
                ga[q][j] = pre_hsig ? av : hsig2(av);
                gf[q][j] = bf2_unpack(reinterpret_cast<const uint32_t*>(&qf[q])[j]);
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(2 * bi - 1 + (k >> 1), Hg, H, h0, h1, lh);   // (h0, h1) == (ra, rb) whenever lh != 0
            bilinear_src(2 * bj - 1 + (k & 1), Wg, W, w0, w1, lw);
            const float w00 = (1.f - lh) * (1.f - lw), w01 = (1.f - lh) * lw, w10 = lh * (1.f - lw), w11 = lh * lw;
            const float2 v00 = make_float2(w00, w00), v01 = make_float2(w01, w01), v10 = make_float2(w10, w10), v11 = make_float2(w11, w11);
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 sig = __ffma2_rn(v11, ga[3][j], __ffma2_rn(v10, ga[2][j], __ffma2_rn(v01, ga[1][j], __fmul2_rn(v00, ga[0][j]))));
                const float2 g = __ffma2_rn(v11, gf[3][j], __ffma2_rn(v10, gf[2][j], __ffma2_rn(v01, gf[1][j], __fmul2_rn(v00, gf[0][j]))));
                const float2 r = __ffma2_rn(bf2_unpack(reinterpret_cast<const uint32_t*>(&ql[k])[j]), sig, g);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(r.x, r.y);
                ow[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            *reinterpret_cast<uint4*>(y + pix[k] * y_cs + c) = o;
        }
    }
}
}  // namespace mgdt

extern "C" int mgdt_inject(const void* local, int l_cs, const void* gact, int a_cs, const void* gfeat, int f_cs,
                           void* y, int y_cs, int N, int H, int W, int Hg, int Wg, int C, int dtype, void* stream) {
    return mgdt_inject2(local, l_cs, gact, a_cs, gfeat, f_cs, y, y_cs, N, H, W, Hg, Wg, C, 0, dtype, stream);
}

extern "C" int mgdt_inject2(const void* local, int l_cs, const void* gact, int a_cs, const void* gfeat, int f_cs,
                            void* y, int y_cs, int N, int H, int W, int Hg, int Wg, int C, int gact_is_hsig, int dtype, void* stream) {
    MGDT_CHECK(local && gact && gfeat && y, "inject: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0 && C > 0, "inject: bad shape");
    MGDT_CHECK((long long)N * H * W * C < (1LL << 31) && (long long)N * Hg * Wg * C < (1LL << 31),
               "inject: tensor too large for 32-bit indexing");
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && aligned8(local, l_cs, sizeof(T)) && aligned8(gact, a_cs, sizeof(T)) &&
                         aligned8(gfeat, f_cs, sizeof(T)) && aligned8(y, y_cs, sizeof(T));
        MGDT_VEC_SWITCH(vec, V, {
            if (H == 2 * Hg && W == 2 * Wg && Hg > 1 && Wg > 1) {   // exact 2x upsampling: one thread per 2x2 output block
                const unsigned total2 = (unsigned)((long long)N * (Hg + 1) * (Wg + 1) * (C / V));
                if constexpr (sizeof(T) == 2 && V == 8)
                    launch_k(inject2x_bf16_kernel, dim3(ew_grid(total2, 1)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)local, l_cs,
                             (const T*)gact, a_cs, (const T*)gfeat, f_cs, (T*)y, y_cs, Hg, Wg, (unsigned)C, total2, gact_is_hsig);
                else if (gact_is_hsig)
                    return set_error(-ENOTSUP, "inject: a pre-activated gate is implemented for the bf16 exact-2x kernel only");
                else
                    launch_k(inject2x_kernel<T, V>, dim3(ew_grid(total2)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)local, l_cs,
                             (const T*)gact, a_cs, (const T*)gfeat, f_cs, (T*)y, y_cs, Hg, Wg, (unsigned)C, total2);
                MGDT_LAUNCH_CHECK("inject");
                return 0;
            }
            if (gact_is_hsig) return set_error(-ENOTSUP, "inject: a pre-activated gate is implemented for the bf16 exact-2x kernel only");
            const unsigned total = (unsigned)((long long)N * H * W * (C / V));
            launch_k(inject_kernel<T, V>, dim3(ew_grid(total)), dim3(EW_THREADS), 0, (cudaStream_t)stream, (const T*)local, l_cs, (const T*)gact, a_cs, (const T*)gfeat, f_cs, (T*)y, y_cs, H, W, Hg, Wg,
                (unsigned)C, total);
        });
    });
    MGDT_LAUNCH_CHECK("inject");
    return 0;
}

extern "C" int mgdt_preprocess(const void* src, int src_is_u8, void* y, int y_cs, int N, int C, int H, int W, int dtype,
                               void* stream) {
    MGDT_CHECK(src && y && N > 0 && C > 0 && H > 0 && W > 0 && y_cs >= C, "preprocess: bad args");
    MGDT_CHECK((long long)N * H * W * C < (1LL << 31), "preprocess: tensor too large for 32-bit indexing");
    const unsigned npix = (unsigned)((long long)N * H * W);
    const unsigned HW = (unsigned)(H * W);
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_DTYPE_SWITCH(dtype, T, {
        if (src_is_u8)
            launch_k(preprocess_kernel<uint8_t, T>, dim3(ew_grid(npix)), dim3(EW_THREADS), 0, s, (const uint8_t*)src, (T*)y, y_cs, C, HW,
                                                                               255.0f, npix);
        else
            launch_k(preprocess_kernel<float, T>, dim3(ew_grid(npix)), dim3(EW_THREADS), 0, s, (const float*)src, (T*)y, y_cs, C, HW,
                                                                             1.0f, npix);
    });
    MGDT_LAUNCH_CHECK("preprocess");
    return 0;
}
