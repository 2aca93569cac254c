// Per-(n,c) reductions over the image and the tiny gate / normalisation-finalize kernels that
// consume them.  All reductions are deterministic (two-stage, no float atomics).
#include "common.cuh"

namespace mgdt {

// ------------------------------------------------------------------ chan_stats
// Stage 1: grid (chunks, N).  A block walks a strip of pixels; thread = (pixel group, channel vector
// of V channels); partial sums go to part[n][chunk][Q][C] and partsq[n][chunk][C].
constexpr int CS_THREADS = 256;

template <typename T, int V> struct StatLd;
template <typename T> struct StatLd<T, 1> {
    static __device__ __forceinline__ void ld(const T* p, float* f) { f[0] = ldf(p); }
};
template <> struct StatLd<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
    }
};
template <> struct StatLd<float, 8> {
    static __device__ __forceinline__ void ld(const float* p, float* f) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
};

template <typename T, int V, int Q>
__global__ void __launch_bounds__(CS_THREADS) chan_stats_partial(const T* __restrict__ x, int x_cs, int H, int W, int C,
                                                                 int pix_per_chunk, float* __restrict__ part,
                                                                 float* __restrict__ partsq) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[CS_THREADS * V];  // [PG][Cw * V]
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x;
    const int CV = C / V;
    int Cw = 1;
    while (Cw < CV && Cw < CS_THREADS) Cw <<= 1;  // channel-vector lanes per pass (pow2 <= 256)
    const int PG = CS_THREADS / Cw;
    const int cl = threadIdx.x % Cw, pg = threadIdx.x / Cw;
    const int HW = H * W;
    const int p_begin = chunk * pix_per_chunk;
    const int p_end = min(HW, p_begin + pix_per_chunk);
    // adaptive_avg_pool2d(2) windows: [floor(i*H/2), ceil((i+1)*H/2))
    const int h_top_end = (H + 1) / 2, h_bot_begin = H / 2;
    const int w_left_end = (W + 1) / 2, w_right_begin = W / 2;
    const T* xn = x + (size_t)n * HW * x_cs;

    for (int cb = 0; cb < CV; cb += Cw) {
        const int cv = cb + cl;
        float s[Q][V];
        float sq[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            sq[j] = 0.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) s[q][j] = 0.f;
        }
        if (cv < CV) {
#pragma unroll 4
            for (int pidx = p_begin + pg; pidx < p_end; pidx += PG) {
                float v[V];
                StatLd<T, V>::ld(xn + (size_t)pidx * x_cs + cv * V, v);
                bool m[4] = {false, false, false, false};
                if (Q == 5) {
                    const int h = pidx / W, w = pidx - h * W;
                    const bool top = h < h_top_end, bot = h >= h_bot_begin;
                    const bool left = w < w_left_end, right = w >= w_right_begin;
                    m[0] = top && left; m[1] = top && right; m[2] = bot && left; m[3] = bot && right;
                }
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    s[0][j] += v[j];
                    sq[j] += v[j] * v[j];
                    if (Q == 5) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (m[q]) s[1 + q][j] += v[j];
                    }
                }
            }
        }
        // reduce over pixel groups, one statistic at a time: xor-shuffles inside the warp (lanes that share a channel
        // lane are Cw apart), then at most 8 partial rows through shared memory (fixed order -> deterministic)
        const int lane = threadIdx.x & 31;
        const int rows = Cw <= 32 ? CS_THREADS / 32 : PG;
        const int row = Cw <= 32 ? (threadIdx.x >> 5) : pg;
#pragma unroll
        for (int q = 0; q <= Q; ++q) {
            if (q == Q && !partsq) break;
            float t[V];
#pragma unroll
            for (int j = 0; j < V; ++j) t[j] = (q < Q) ? s[q < Q ? q : 0][j] : sq[j];
            for (int off = 16; off >= Cw; off >>= 1) {
#pragma unroll
                for (int j = 0; j < V; ++j) t[j] += __shfl_xor_sync(0xffffffffu, t[j], off);
            }
            if (Cw >= 32 || lane < Cw) {
#pragma unroll
                for (int j = 0; j < V; ++j) sm[(row * Cw + cl) * V + j] = t[j];
            }
            __syncthreads();
            if ((int)threadIdx.x < Cw && cb + (int)threadIdx.x < CV) {
                const int cvo = cb + (int)threadIdx.x;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float tot = 0.f;
                    for (int g = 0; g < rows; ++g) tot += sm[(g * Cw + (int)threadIdx.x) * V + j];
                    if (q < Q)
                        part[(((size_t)n * nchunks + chunk) * Q + q) * C + cvo * V + j] = tot;
                    else
                        partsq[((size_t)n * nchunks + chunk) * C + cvo * V + j] = tot;
                }
            }
            __syncthreads();
        }
    }
}

__global__ void chan_stats_final(const float* __restrict__ part, const float* __restrict__ partsq, int nchunks, int Q,
                                 int C, float* __restrict__ out_sum, float* __restrict__ out_sumsq, int N) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * Q * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total + (long long)N * C;
         i += (long long)gridDim.x * blockDim.x) {
        if (i < total) {
            const int c = (int)(i % C);
            const int q = (int)((i / C) % Q);
            const int n = (int)(i / ((long long)C * Q));
            float t = 0.f;
            for (int k = 0; k < nchunks; ++k) t += part[(((long long)n * nchunks + k) * Q + q) * C + c];
            out_sum[i] = t;
        } else if (out_sumsq) {
            const long long j = i - total;
            const int c = (int)(j % C);
            const int n = (int)(j / C);
            float t = 0.f;
            for (int k = 0; k < nchunks; ++k) t += partsq[((long long)n * nchunks + k) * C + c];
            out_sumsq[j] = t;
        }
    }
}

static int stats_chunks(int N, int H, int W) {
    // enough blocks to fill 148 SMs a few times, at least 64 pixels per chunk
    const int HW = H * W;
    int want = (148 * 4 + N - 1) / N;
    int maxc = (HW + 63) / 64;
    int c = want < maxc ? want : maxc;
    return c < 1 ? 1 : c;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" size_t mgdt_chan_stats_ws_bytes(int N, int H, int W, int C, int quads) {
    const int Q = quads ? 5 : 1;
    const size_t nch = (size_t)stats_chunks(N, H, W);
    return sizeof(float) * (size_t)N * nch * (size_t)(Q + 1) * (size_t)C;
}

extern "C" int mgdt_chan_stats(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum,
                               float* out_sumsq, void* ws, size_t ws_bytes, int dtype, void* stream) {
    MGDT_CHECK(x && out_sum && ws, "chan_stats: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && x_cs >= C, "chan_stats: bad shape");
    MGDT_CHECK(ws_bytes >= mgdt_chan_stats_ws_bytes(N, H, W, C, quads), "chan_stats: workspace too small");
    const int Q = quads ? 5 : 1;
    const int nch = stats_chunks(N, H, W);
    const int ppc = cdiv(H * W, nch);
    float* part = (float*)ws;
    float* partsq = part + (size_t)N * nch * Q * C;
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && (x_cs % 8) == 0 && (((uintptr_t)x) % (8 * sizeof(T))) == 0;
        float* psq = out_sumsq ? partsq : nullptr;
        if (vec && quads) launch_k(chan_stats_partial<T, 8, 5>, dim3(dim3(nch, N)), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, ppc, part, psq);
        else if (vec) launch_k(chan_stats_partial<T, 8, 1>, dim3(dim3(nch, N)), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, ppc, part, psq);
        else if (quads) launch_k(chan_stats_partial<T, 1, 5>, dim3(dim3(nch, N)), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, ppc, part, psq);
        else launch_k(chan_stats_partial<T, 1, 1>, dim3(dim3(nch, N)), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, ppc, part, psq);
    });
    MGDT_LAUNCH_CHECK("chan_stats_partial");
    const long long total = (long long)N * (Q + 1) * C;
    launch_k(chan_stats_final, dim3(cdiv(total, 256)), dim3(256), 0, s, part, partsq, nch, Q, C, out_sum, out_sumsq, N);
    MGDT_LAUNCH_CHECK("chan_stats_final");
    return 0;
}

// ------------------------------------------------------------------ MSPA gate
// One block per image, thread per (group g, channel c) pair for the MLP output.
namespace mgdt {
__global__ void mspa_gate_kernel(const float* __restrict__ stats, int H, int W, int C, int G, int softmax,
                                 const float* __restrict__ w1, const float* __restrict__ b1,
                                 const float* __restrict__ w2, const float* __restrict__ b2, int hidden,
                                 float* __restrict__ scale) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];  // feat[G][5*ow] | hid[G][hidden] | gate[G][ow]
    const int n = blockIdx.x;
    const int ow = C / G;
    float* feat = sm;
    float* hid = feat + G * 5 * ow;
    float* gate = hid + G * hidden;
    const float* st = stats + (long long)n * 5 * C;
    // window sizes of adaptive_avg_pool2d(2)
    const int hs[2] = {(H + 1) / 2, H - H / 2};
    const int ws[2] = {(W + 1) / 2, W - W / 2};
    for (int i = threadIdx.x; i < G * ow; i += blockDim.x) {
        const int g = i / ow, c = i % ow;
        const int ch = g * ow + c;
        feat[g * 5 * ow + c] = st[0 * C + ch] / (float)(H * W);
        for (int q = 0; q < 4; ++q)
            feat[g * 5 * ow + ow + c * 4 + q] = st[(1 + q) * C + ch] / (float)(hs[q >> 1] * ws[q & 1]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * hidden; i += blockDim.x) {
        const int g = i / hidden, j = i % hidden;
        float a = b1[j];
        const float* wr = w1 + (long long)j * 5 * ow;
        const float* f = feat + g * 5 * ow;
        for (int k = 0; k < 5 * ow; ++k) a = fmaf(wr[k], f[k], a);
        hid[i] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * ow; i += blockDim.x) {
        const int g = i / ow, c = i % ow;
        float a = b2[c];
        for (int k = 0; k < hidden; ++k) a = fmaf(w2[c * hidden + k], hid[g * hidden + k], a);
        gate[i] = sigmoidf_(a);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ow; c += blockDim.x) {
        if (!softmax) {
            for (int g = 0; g < G; ++g) scale[(long long)n * C + g * ow + c] = gate[g * ow + c];
            continue;
        }
        float mx = -1e30f, den = 0.f;
        for (int g = 0; g < G; ++g) mx = fmaxf(mx, gate[g * ow + c]);
        for (int g = 0; g < G; ++g) den += expf(gate[g * ow + c] - mx);
        for (int g = 0; g < G; ++g) scale[(long long)n * C + g * ow + c] = expf(gate[g * ow + c] - mx) / den;
    }
}

__global__ void grn_scale_kernel(const float* __restrict__ sumsq, const float* __restrict__ gamma, int C,
                                 float* __restrict__ scale) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    const int n = blockIdx.x;
    float part = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf(sumsq[(long long)n * C + c]);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
    const float denom = tot / (float)C + 1e-6f;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        scale[(long long)n * C + c] = 1.0f + gamma[c] * (sqrtf(sumsq[(long long)n * C + c]) / denom);
}

__global__ void gn_affine_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, int C, int groups,
                                 int hw, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ a, float* __restrict__ b) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x;
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < cpg; ++k) {
            s += (double)sum[(long long)n * C + g * cpg + k];
            q += (double)sumsq[(long long)n * C + g * cpg + k];
        }
        const double cnt = (double)cpg * (double)hw;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        for (int k = 0; k < cpg; ++k) {
            const int c = g * cpg + k;
            const float aa = gamma[c] * rstd;
            a[(long long)n * C + c] = aa;
            b[(long long)n * C + c] = beta[c] - (float)mean * aa;
        }
    }
}

__global__ void td_attn_kernel(const float* __restrict__ sum, int N, int C, int hw, int hidden, int stacked,
                               const float* __restrict__ w1, const float* __restrict__ b1,
                               const float* __restrict__ w2, const float* __restrict__ b2,
                               float* __restrict__ in_scale) {
    pdl_trigger();
    pdl_wait();
    // grid (N, ndec): blockIdx.y selects the decomposition (0 = cls, 1 = reg); weights are packed per
    // decomposition back to back.
    extern __shared__ float sm[];  // mean[C] | hid[hidden] | att[stacked]
    const int n = blockIdx.x, which = blockIdx.y;
    float* mean = sm;
    float* hid = mean + C;
    float* att = hid + hidden;
    const float* W1 = w1 + (long long)which * hidden * C;
    const float* B1 = b1 + which * hidden;
    const float* W2 = w2 + (long long)which * stacked * hidden;
    const float* B2 = b2 + which * stacked;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mean[c] = sum[(long long)n * C + c] / (float)hw;
    __syncthreads();
    for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
        float a = B1[j];
        for (int c = 0; c < C; ++c) a = fmaf(W1[j * C + c], mean[c], a);
        hid[j] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < stacked; s += blockDim.x) {
        float a = B2[s];
        for (int j = 0; j < hidden; ++j) a = fmaf(W2[s * hidden + j], hid[j], a);
        att[s] = sigmoidf_(a);
    }
    __syncthreads();
    const int fc = C / stacked;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        in_scale[((long long)which * N + n) * C + c] = att[c / fc];
}
}  // namespace mgdt

extern "C" int mgdt_mspa_gate(const float* stats, int N, int H, int W, int C, int groups, int softmax,
                              const float* fc1_w, const float* fc1_b, const float* fc2_w, const float* fc2_b,
                              int hidden, float* scale, void* stream) {
    MGDT_CHECK(stats && fc1_w && fc1_b && fc2_w && fc2_b && scale, "mspa_gate: null pointer");
    MGDT_CHECK(N > 0 && C > 0 && groups > 0 && C % groups == 0 && hidden > 0, "mspa_gate: bad shape C=%d groups=%d hidden=%d",
               C, groups, hidden);
    const int ow = C / groups;
    const size_t smem = sizeof(float) * (size_t)groups * (5 * ow + hidden + ow);
    MGDT_CHECK(smem <= 48 * 1024, "mspa_gate: C=%d too large", C);
    launch_k(mspa_gate_kernel, dim3(N), dim3(128), smem, (cudaStream_t)stream, stats, H, W, C, groups, softmax, fc1_w, fc1_b, fc2_w, fc2_b,
                                                             hidden, scale);
    MGDT_LAUNCH_CHECK("mspa_gate");
    return 0;
}

extern "C" int mgdt_grn_scale(const float* sumsq, const float* gamma, int N, int C, float* scale, void* stream) {
    MGDT_CHECK(sumsq && gamma && scale && N > 0 && C > 0, "grn_scale: bad args");
    launch_k(grn_scale_kernel, dim3(N), dim3(128), 0, (cudaStream_t)stream, sumsq, gamma, C, scale);
    MGDT_LAUNCH_CHECK("grn_scale");
    return 0;
}

extern "C" int mgdt_gn_affine(const float* sum, const float* sumsq, int N, int C, int groups, int hw, float eps,
                              const float* gamma, const float* beta, float* a, float* b, void* stream) {
    MGDT_CHECK(sum && sumsq && gamma && beta && a && b, "gn_affine: null pointer");
    MGDT_CHECK(N > 0 && groups > 0 && C % groups == 0 && hw > 0, "gn_affine: bad shape");
    launch_k(gn_affine_kernel, dim3(N), dim3(32), 0, (cudaStream_t)stream, sum, sumsq, C, groups, hw, eps, gamma, beta, a, b);
    MGDT_LAUNCH_CHECK("gn_affine");
    return 0;
}

extern "C" int mgdt_td_attn(const float* sum, int N, int C, int hw, int hidden, int stacked, int ndec,
                            const float* la1_w, const float* la1_b, const float* la2_w, const float* la2_b,
                            float* in_scale, void* stream) {
    MGDT_CHECK(sum && la1_w && la1_b && la2_w && la2_b && in_scale, "td_attn: null pointer");
    MGDT_CHECK(N > 0 && C > 0 && stacked > 0 && C % stacked == 0 && hidden > 0 && ndec > 0, "td_attn: bad shape");
    const size_t smem = sizeof(float) * (C + hidden + stacked);
    launch_k(td_attn_kernel, dim3(dim3(N, ndec)), dim3(64), smem, (cudaStream_t)stream, sum, N, C, hw, hidden, stacked, la1_w, la1_b, la2_w,
                                                                   la2_b, in_scale);
    MGDT_LAUNCH_CHECK("td_attn");
    return 0;
}
