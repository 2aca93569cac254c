// Memory-bound NHWC kernels: per-(n,c) affine + activation, resampling into concat slices, the
// SPPF pooling chain, the injection gate and input preprocessing.  Thread <-> (pixel, channel)
// with the channel fastest so that every warp touches contiguous bytes.
#include "common.cuh"

namespace mgdt {

constexpr int EW_THREADS = 256;

static inline int ew_grid(long long total) {
    long long b = (total + EW_THREADS - 1) / EW_THREADS;
    const long long cap = 148LL * 16;  // grid-stride beyond 16 resident CTAs per SM
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ------------------------------------------------------------------ affine + act (+ other)
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) affine_act_kernel(const T* __restrict__ x, int x_cs,
                                                                const float* __restrict__ a,
                                                                const float* __restrict__ b,
                                                                const T* __restrict__ other, int o_cs, int act,
                                                                T* __restrict__ y, int y_cs, long long HW, int C,
                                                                long long total) {
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW;
        float v = ldf(x + pix * x_cs + c);
        if (a) v *= a[n * C + c];
        if (b) v += b[n * C + c];
        v = apply_act(v, act);
        if (other) v += ldf(other + pix * o_cs + c);
        stf(y + pix * y_cs + c, v);
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

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) resample_kernel(const T* __restrict__ x, int x_cs, int Hi, int Wi,
                                                              T* __restrict__ y, int y_cs, int Ho, int Wo, int C,
                                                              int mode, long long total) {
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
        const int c = (int)(i % C);
        long long pix = i / C;
        const int wo = (int)(pix % Wo);
        const int ho = (int)((pix / Wo) % Ho);
        const long long n = pix / ((long long)Wo * Ho);
        const T* xn = x + n * (long long)Hi * Wi * x_cs + c;
        float v;
        if (mode == MGDT_RS_COPY) {
            v = ldf(xn + ((long long)ho * Wi + wo) * x_cs);
        } else if (mode == MGDT_RS_NEAREST) {
            // nn.Upsample(mode='nearest'): src = floor(dst * in/out)
            const int hi = min((int)floorf((float)ho * ((float)Hi / (float)Ho)), Hi - 1);
            const int wi = min((int)floorf((float)wo * ((float)Wi / (float)Wo)), Wi - 1);
            v = ldf(xn + ((long long)hi * Wi + wi) * x_cs);
        } else if (mode == MGDT_RS_AVGPOOL) {
            // adaptive_avg_pool2d: window [floor(o*in/out), ceil((o+1)*in/out))
            const int h0 = (int)(((long long)ho * Hi) / Ho), h1 = (int)((((long long)ho + 1) * Hi + Ho - 1) / Ho);
            const int w0 = (int)(((long long)wo * Wi) / Wo), w1 = (int)((((long long)wo + 1) * Wi + Wo - 1) / Wo);
            float s = 0.f;
            for (int h = h0; h < h1; ++h)
                for (int w = w0; w < w1; ++w) s += ldf(xn + ((long long)h * Wi + w) * x_cs);
            v = s / (float)((h1 - h0) * (w1 - w0));
        } else {
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(ho, Hi, Ho, h0, h1, lh);
            bilinear_src(wo, Wi, Wo, w0, w1, lw);
            const float v00 = ldf(xn + ((long long)h0 * Wi + w0) * x_cs), v01 = ldf(xn + ((long long)h0 * Wi + w1) * x_cs);
            const float v10 = ldf(xn + ((long long)h1 * Wi + w0) * x_cs), v11 = ldf(xn + ((long long)h1 * Wi + w1) * x_cs);
            v = (1.f - lh) * ((1.f - lw) * v00 + lw * v01) + lh * ((1.f - lw) * v10 + lw * v11);
        }
        stf(y + pix * y_cs + c, v);
    }
}

// ------------------------------------------------------------------ SPPF pooling chain
// maxpool(k,1,k/2) applied 1x/2x/3x == max over (k-1)*j+1 windows, j = 1..3 (padding is -inf).
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) sppf_pool_kernel(const T* __restrict__ x, int x_cs, T* __restrict__ y1,
                                                               T* __restrict__ y2, T* __restrict__ y3, int y_cs, int H,
                                                               int W, int C, int r, long long total) {
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
        const int c = (int)(i % C);
        long long pix = i / C;
        const int w = (int)(pix % W);
        const int h = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const T* xn = x + n * (long long)H * W * x_cs + c;
        float m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        for (int dy = -3 * r; dy <= 3 * r; ++dy) {
            const int hh = h + dy;
            if (hh < 0 || hh >= H) continue;
            const int ady = dy < 0 ? -dy : dy;
            for (int dx = -3 * r; dx <= 3 * r; ++dx) {
                const int ww = w + dx;
                if (ww < 0 || ww >= W) continue;
                const int adx = dx < 0 ? -dx : dx;
                const float v = ldf(xn + ((long long)hh * W + ww) * x_cs);
                const int d = ady > adx ? ady : adx;
                m3 = fmaxf(m3, v);
                if (d <= 2 * r) m2 = fmaxf(m2, v);
                if (d <= r) m1 = fmaxf(m1, v);
            }
        }
        stf(y1 + pix * y_cs + c, m1);
        stf(y2 + pix * y_cs + c, m2);
        stf(y3 + pix * y_cs + c, m3);
    }
}

// ------------------------------------------------------------------ injection gate
template <typename T>
__global__ void __launch_bounds__(EW_THREADS) inject_kernel(const T* __restrict__ local, int l_cs,
                                                            const T* __restrict__ gact, int a_cs,
                                                            const T* __restrict__ gfeat, int f_cs, T* __restrict__ y,
                                                            int y_cs, int H, int W, int Hg, int Wg, int C,
                                                            long long total) {
    const bool pool = H < Hg;
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
        const int c = (int)(i % C);
        long long pix = i / C;
        const int w = (int)(pix % W);
        const int h = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const T* an = gact + n * (long long)Hg * Wg * a_cs + c;
        const T* fn = gfeat + n * (long long)Hg * Wg * f_cs + c;
        float sig, gf;
        if (pool) {
            const int h0 = (int)(((long long)h * Hg) / H), h1 = (int)((((long long)h + 1) * Hg + H - 1) / H);
            const int w0 = (int)(((long long)w * Wg) / W), w1 = (int)((((long long)w + 1) * Wg + W - 1) / W);
            float sa = 0.f, sf = 0.f;
            for (int hh = h0; hh < h1; ++hh)
                for (int ww = w0; ww < w1; ++ww) {
                    sa += ldf(an + ((long long)hh * Wg + ww) * a_cs);
                    sf += ldf(fn + ((long long)hh * Wg + ww) * f_cs);
                }
            const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
            sig = sa * inv;
            gf = sf * inv;
        } else {
            int h0, h1, w0, w1;
            float lh, lw;
            bilinear_src(h, Hg, H, h0, h1, lh);
            bilinear_src(w, Wg, W, w0, w1, lw);
            const long long o00 = (long long)h0 * Wg + w0, o01 = (long long)h0 * Wg + w1;
            const long long o10 = (long long)h1 * Wg + w0, o11 = (long long)h1 * Wg + w1;
            const float a00 = apply_act(ldf(an + o00 * a_cs), MGDT_ACT_HSIGMOID), a01 = apply_act(ldf(an + o01 * a_cs), MGDT_ACT_HSIGMOID);
            const float a10 = apply_act(ldf(an + o10 * a_cs), MGDT_ACT_HSIGMOID), a11 = apply_act(ldf(an + o11 * a_cs), MGDT_ACT_HSIGMOID);
            sig = (1.f - lh) * ((1.f - lw) * a00 + lw * a01) + lh * ((1.f - lw) * a10 + lw * a11);
            const float f00 = ldf(fn + o00 * f_cs), f01 = ldf(fn + o01 * f_cs);
            const float f10 = ldf(fn + o10 * f_cs), f11 = ldf(fn + o11 * f_cs);
            gf = (1.f - lh) * ((1.f - lw) * f00 + lw * f01) + lh * ((1.f - lw) * f10 + lw * f11);
        }
        const float l = ldf(local + pix * l_cs + c);
        stf(y + pix * y_cs + c, l * sig + gf);
    }
}

// ------------------------------------------------------------------ preprocess
template <typename S, typename T>
__global__ void __launch_bounds__(EW_THREADS) preprocess_kernel(const S* __restrict__ src, T* __restrict__ y, int y_cs,
                                                                int C, long long HW, float div, long long total) {
    // output-major mapping (channel fastest) -> coalesced NHWC stores; the NCHW reads of one warp hit
    // C planes at consecutive pixels, which L1/L2 merge.
    for (long long i = blockIdx.x * (long long)EW_THREADS + threadIdx.x; i < total;
         i += (long long)gridDim.x * EW_THREADS) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW, p = pix - n * HW;
        const float v = (float)src[(n * C + c) * HW + p] / div;  // `im /= 255` (predictor.py:129)
        stf(y + pix * y_cs + c, v);
    }
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_affine_act(const void* x, int x_cs, const float* a, const float* b, const void* other, int o_cs,
                               int act, void* y, int y_cs, int N, int H, int W, int C, int dtype, void* stream) {
    MGDT_CHECK(x && y, "affine_act: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && x_cs >= C && y_cs >= C, "affine_act: bad shape");
    const long long total = (long long)N * H * W * C;
    MGDT_DTYPE_SWITCH(dtype, T, {
        affine_act_kernel<T><<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
            (const T*)x, x_cs, a, b, (const T*)other, o_cs, act, (T*)y, y_cs, (long long)H * W, C, total);
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
    const long long total = (long long)N * Ho * Wo * C;
    MGDT_DTYPE_SWITCH(dtype, T, {
        resample_kernel<T><<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const T*)x, x_cs, Hi, Wi, (T*)y,
                                                                                     y_cs, Ho, Wo, C, mode, total);
    });
    MGDT_LAUNCH_CHECK("resample");
    return 0;
}

extern "C" int mgdt_sppf_pool(const void* x, int x_cs, void* y1, void* y2, void* y3, int y_cs, int N, int H, int W,
                              int C, int k, int dtype, void* stream) {
    MGDT_CHECK(x && y1 && y2 && y3, "sppf_pool: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && (k & 1) && k >= 1, "sppf_pool: bad shape/k");
    const long long total = (long long)N * H * W * C;
    MGDT_DTYPE_SWITCH(dtype, T, {
        sppf_pool_kernel<T><<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
            (const T*)x, x_cs, (T*)y1, (T*)y2, (T*)y3, y_cs, H, W, C, k / 2, total);
    });
    MGDT_LAUNCH_CHECK("sppf_pool");
    return 0;
}

extern "C" int mgdt_inject(const void* local, int l_cs, const void* gact, int a_cs, const void* gfeat, int f_cs,
                           void* y, int y_cs, int N, int H, int W, int Hg, int Wg, int C, int dtype, void* stream) {
    MGDT_CHECK(local && gact && gfeat && y, "inject: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && Hg > 0 && Wg > 0 && C > 0, "inject: bad shape");
    const long long total = (long long)N * H * W * C;
    MGDT_DTYPE_SWITCH(dtype, T, {
        inject_kernel<T><<<ew_grid(total), EW_THREADS, 0, (cudaStream_t)stream>>>(
            (const T*)local, l_cs, (const T*)gact, a_cs, (const T*)gfeat, f_cs, (T*)y, y_cs, H, W, Hg, Wg, C, total);
    });
    MGDT_LAUNCH_CHECK("inject");
    return 0;
}

extern "C" int mgdt_preprocess(const void* src, int src_is_u8, void* y, int y_cs, int N, int C, int H, int W, int dtype,
                               void* stream) {
    MGDT_CHECK(src && y && N > 0 && C > 0 && H > 0 && W > 0 && y_cs >= C, "preprocess: bad args");
    const long long total = (long long)N * H * W * C;
    const long long HW = (long long)H * W;
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_DTYPE_SWITCH(dtype, T, {
        if (src_is_u8)
            preprocess_kernel<uint8_t, T><<<ew_grid(total), EW_THREADS, 0, s>>>((const uint8_t*)src, (T*)y, y_cs, C, HW,
                                                                                255.0f, total);
        else
            preprocess_kernel<float, T><<<ew_grid(total), EW_THREADS, 0, s>>>((const float*)src, (T*)y, y_cs, C, HW,
                                                                              1.0f, total);
    });
    MGDT_LAUNCH_CHECK("preprocess");
    return 0;
}
