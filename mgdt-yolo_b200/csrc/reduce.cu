// Per-(n,c) reductions over the image and the tiny gate / normalisation-finalize kernels that
// consume them.  All reductions are deterministic (two-stage, no float atomics).
#include "common.cuh"

#include <algorithm>

namespace mgdt {

// ------------------------------------------------------------------ chan_stats
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

// ------------------------------------------------------------------ per-image finalisers
// Bodies of the tiny per-image kernels (SPR gate MLP, GRN scale, GroupNorm affine) as device functions over one image:
// they run either as their own one-block-per-image kernels or inside the last block of chan_stats (mgdt_stats_fin),
// which saves a launch per use.  `sm` is scratch shared memory, blockDim-stride loops throughout.
struct StatsFin {
    int kind;                  // 0 none, 1 SPR gate, 2 GRN scale, 3 GroupNorm affine
    const float *p0, *p1, *p2, *p3;
    int i0, i1, i2;
    float f0;
    float *o0, *o1;
};

__device__ __forceinline__ void mspa_gate_body(int n, const float* st, int H, int W, int C, int G, int softmax,
                                               const float* __restrict__ w1, const float* __restrict__ b1,
                                               const float* __restrict__ w2, const float* __restrict__ b2, int hidden,
                                               float* __restrict__ scale, float* sm) {
    const int ow = C / G;
    float* feat = sm;                    // feat[G][5*ow] | hid[G][hidden] | gate[G][ow]
    float* hid = feat + G * 5 * ow;
    float* gate = hid + G * hidden;
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
    // fc1: one warp per output, lanes stride the 5*ow inputs (coalesced weight reads), fixed-order butterfly sum
    for (int i = threadIdx.x >> 5; i < G * hidden; i += (int)(blockDim.x >> 5)) {
        const int g = i / hidden, j = i % hidden;
        const float* wr = w1 + (long long)j * 5 * ow;
        const float* f = feat + g * 5 * ow;
        float a = 0.f;
        for (int k = threadIdx.x & 31; k < 5 * ow; k += 32) a = fmaf(wr[k], f[k], a);
        a = warp_sum(a);
        if ((threadIdx.x & 31) == 0) hid[i] = fmaxf(a + b1[j], 0.f);
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

__device__ __forceinline__ void grn_scale_body(int n, const float* sumsq_n, const float* __restrict__ gamma, int C,
                                               float* __restrict__ scale, float* red) {
    float part = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) part += sqrtf(sumsq_n[c]);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const float denom = tot / (float)C + 1e-6f;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        scale[(long long)n * C + c] = 1.0f + gamma[c] * (sqrtf(sumsq_n[c]) / denom);
}

__device__ __forceinline__ void gn_affine_body(int n, const float* sum_n, const float* sumsq_n, int C, int groups, int hw,
                                               float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                               float* __restrict__ a, float* __restrict__ b) {
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < cpg; ++k) {
            s += (double)sum_n[g * cpg + k];
            q += (double)sumsq_n[g * cpg + k];
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

// Stage 1: grid (chunks, N, rects).  rects = 1 (whole image) or, with quads, the four adaptive_avg_pool2d(2) windows
// [floor(i*H/2), ceil((i+1)*H/2)) (+ a fifth whole-image rect when H or W is odd and the windows overlap; for even
// sizes the total is the sum of the four).  A block walks a strip of its rectangle's pixels; thread = (pixel group,
// channel vector of V channels) with 8 (+8) accumulators, so the kernel stays at full occupancy.  Partial sums go to
// part[n][chunk][rect][C] and partsq[n][chunk][rect][C]; the LAST block of an image to finish (atomic ticket) reduces
// them in chunk order (fixed order -> deterministic) into out_sum[n][Q][C] / out_sumsq[n][C].
template <typename T, int V>
__global__ void __launch_bounds__(CS_THREADS) chan_stats_partial(const T* __restrict__ x, int x_cs, int H, int W, int C, int Q,
                                                                 float* __restrict__ part, float* __restrict__ partsq,
                                                                 int* __restrict__ counters, float* __restrict__ out_sum,
                                                                 float* __restrict__ out_sumsq, StatsFin fin) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[CS_THREADS * V];  // [rows][Cw * V]
    const int n = blockIdx.y, chunk = blockIdx.x, nchunks = gridDim.x, rect = blockIdx.z, nrect = gridDim.z;
    const int CV = C / V;
    int Cw = 1;
    while (Cw < CV && Cw < CS_THREADS) Cw <<= 1;  // channel-vector lanes per pass (pow2 <= 256)
    const int PG = CS_THREADS / Cw;
    const int cl = threadIdx.x % Cw, pg = threadIdx.x / Cw;
    // rectangle of this block
    int rh0 = 0, rh1 = H, rw0 = 0, rw1 = W;
    if (Q == 5 && rect < 4) {
        if ((rect >> 1) == 0) rh1 = (H + 1) / 2; else rh0 = H / 2;
        if ((rect & 1) == 0) rw1 = (W + 1) / 2; else rw0 = W / 2;
    }
    const int rw = rw1 - rw0, npx = (rh1 - rh0) * rw;
    const int ppc = (npx + nchunks - 1) / nchunks;
    const int p_begin = chunk * ppc, p_end = min(npx, p_begin + ppc);
    // the sum of squares is only wanted over the whole image: every window when they partition it, else the full rect
    const bool want_sq = partsq != nullptr && (nrect != 5 || rect == 4);
    const T* xn = x + (size_t)n * H * W * x_cs;
    const int lane = threadIdx.x & 31;
    const int rows = Cw <= 32 ? CS_THREADS / 32 : PG;
    const int row = Cw <= 32 ? (threadIdx.x >> 5) : pg;

    for (int cb = 0; cb < CV; cb += Cw) {
        const int cv = cb + cl;
        float s[V], sq[V];
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = sq[j] = 0.f;
        if (cv < CV) {
#pragma unroll 4
            for (int i = p_begin + pg; i < p_end; i += PG) {
                const int hh = i / rw, ww = i - hh * rw;
                float v[V];
                StatLd<T, V>::ld(xn + (size_t)((rh0 + hh) * W + rw0 + ww) * x_cs + cv * V, v);
#pragma unroll
                for (int j = 0; j < V; ++j) { s[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); }
            }
        }
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            if (st == 1 && !want_sq) break;
            float t[V];
#pragma unroll
            for (int j = 0; j < V; ++j) t[j] = st ? sq[j] : s[j];
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
                    float* dst = st == 0 ? part : partsq;
                    dst[(((size_t)n * nchunks + chunk) * nrect + rect) * C + cvo * V + j] = tot;
                }
            }
            __syncthreads();
        }
    }
    // finalisation by the last block of the image
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&counters[n], 1) == nchunks * nrect - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // four lanes per output (chunks k = sub, sub + 4, ...), combined by a fixed xor butterfly: deterministic, and the
    // serial tail of the kernel (one block per image) is four times shorter
    const int sub = threadIdx.x & 3;
    const int n_out = Q * C + (partsq ? C : 0);
    for (int i = threadIdx.x >> 2; i < (n_out + 7) / 8 * 8; i += CS_THREADS / 4) {   // whole warps stay in the loop together
        const bool valid = i < n_out;
        const bool is_sq = i >= Q * C;
        const int q = is_sq ? 0 : i / C, c = is_sq ? i - Q * C : i - q * C;
        const float* src = is_sq ? partsq : part;
        float t = 0.f;
        if (valid) {
            if (q == 0 && nrect == 4) {
                // even H and W: the windows partition the image; total = (q00 + q01) + (q10 + q11)
                float tq[4] = {0.f, 0.f, 0.f, 0.f};
                for (int k = sub; k < nchunks; k += 4)
#pragma unroll
                    for (int r = 0; r < 4; ++r) tq[r] += __ldcg(&src[(((size_t)n * nchunks + k) * nrect + r) * C + c]);
                t = (tq[0] + tq[1]) + (tq[2] + tq[3]);
            } else {
                const int r = nrect == 1 ? 0 : (q == 0 ? 4 : q - 1);
#pragma unroll 2
                for (int k = sub; k < nchunks; k += 4) t += __ldcg(&src[(((size_t)n * nchunks + k) * nrect + r) * C + c]);
            }
        }
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        if (valid && sub == 0) {
            if (is_sq) out_sumsq[(size_t)n * C + c] = t;
            else out_sum[(size_t)n * Q * C + i] = t;
        }
    }
    if (threadIdx.x == 0) counters[n] = 0;
    if (fin.kind) {
        // the image's statistics are complete: finish the consumer's per-image computation here (one launch saved)
        __syncthreads();
        if (fin.kind == 1)
            mspa_gate_body(n, out_sum + (size_t)n * Q * C, H, W, C, fin.i0, fin.i1, fin.p0, fin.p1, fin.p2, fin.p3, fin.i2,
                           fin.o0, sm);
        else if (fin.kind == 2)
            grn_scale_body(n, out_sumsq + (size_t)n * C, fin.p0, C, fin.o0, sm);
        else if (fin.kind == 3)
            gn_affine_body(n, out_sum + (size_t)n * Q * C, out_sumsq + (size_t)n * C, C, fin.i0, H * W, fin.f0, fin.p0, fin.p1,
                           fin.o0, fin.o1);
    }
}

static int stats_rects(int H, int W, int quads) { return quads ? ((H % 2 == 0 && W % 2 == 0) ? 4 : 5) : 1; }

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
    const size_t nch = (size_t)stats_chunks(N, H, W), nrect = (size_t)stats_rects(H, W, quads);
    return sizeof(float) * (size_t)N * nch * nrect * 2 * (size_t)C;
}

static int chan_stats_impl(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum, float* out_sumsq,
                           void* ws, size_t ws_bytes, int32_t* counters, StatsFin fin, int dtype, void* stream) {
    MGDT_CHECK(x && out_sum && ws && counters, "chan_stats: null pointer");
    MGDT_CHECK(N > 0 && H > 0 && W > 0 && C > 0 && x_cs >= C, "chan_stats: bad shape");
    MGDT_CHECK(ws_bytes >= mgdt_chan_stats_ws_bytes(N, H, W, C, quads), "chan_stats: workspace too small");
    const int Q = quads ? 5 : 1;
    const int nrect = stats_rects(H, W, quads);
    // chunks per rectangle (the workspace is sized for stats_chunks() of them; the four windows share that budget)
    const int nch = std::max(1, stats_chunks(N, H, W) / (nrect >= 4 ? 4 : 1));
    float* part = (float*)ws;
    float* partsq = part + (size_t)N * nch * nrect * C;
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_DTYPE_SWITCH(dtype, T, {
        const bool vec = C % 8 == 0 && (x_cs % 8) == 0 && (((uintptr_t)x) % (8 * sizeof(T))) == 0;
        float* psq = out_sumsq ? partsq : nullptr;
        if (vec) launch_k(chan_stats_partial<T, 8>, dim3(nch, N, nrect), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, Q, part, psq, (int*)counters, out_sum, out_sumsq, fin);
        else launch_k(chan_stats_partial<T, 1>, dim3(nch, N, nrect), dim3(CS_THREADS), 0, s, (const T*)x, x_cs, H, W, C, Q, part, psq, (int*)counters, out_sum, out_sumsq, fin);
    });
    MGDT_LAUNCH_CHECK("chan_stats");
    return 0;
}

extern "C" int mgdt_chan_stats(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum,
                               float* out_sumsq, void* ws, size_t ws_bytes, int32_t* counters, int dtype, void* stream) {
    StatsFin fin{};
    return chan_stats_impl(x, x_cs, N, H, W, C, quads, out_sum, out_sumsq, ws, ws_bytes, counters, fin, dtype, stream);
}

extern "C" int mgdt_chan_stats_fin(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum,
                                   float* out_sumsq, void* ws, size_t ws_bytes, int32_t* counters,
                                   const mgdt_stats_fin* f, int dtype, void* stream) {
    MGDT_CHECK(f, "chan_stats_fin: null finaliser");
    StatsFin fin{};
    fin.kind = f->kind; fin.p0 = f->p0; fin.p1 = f->p1; fin.p2 = f->p2; fin.p3 = f->p3;
    fin.i0 = f->i0; fin.i1 = f->i1; fin.i2 = f->i2; fin.f0 = f->f0; fin.o0 = f->o0; fin.o1 = f->o1;
    const bool vec = C % 8 == 0 && (x_cs % 8) == 0 && (((uintptr_t)x) % (8 * (dtype == MGDT_F32 ? 4 : 2))) == 0;
    const int smem_floats = CS_THREADS * (vec ? 8 : 1);
    if (fin.kind == MGDT_FIN_GATE) {
        MGDT_CHECK(quads && fin.p0 && fin.p1 && fin.p2 && fin.p3 && fin.o0 && fin.i0 > 0 && C % fin.i0 == 0 && fin.i2 > 0,
                   "chan_stats_fin: bad SPR gate arguments");
        const int ow = C / fin.i0;
        MGDT_CHECK(fin.i0 * 5 * ow + fin.i0 * fin.i2 + fin.i0 * ow <= smem_floats,
                   "chan_stats_fin: gate scratch does not fit (C=%d): use mgdt_mspa_gate", C);
    } else if (fin.kind == MGDT_FIN_GRN) {
        MGDT_CHECK(out_sumsq && fin.p0 && fin.o0, "chan_stats_fin: bad GRN arguments");
    } else if (fin.kind == MGDT_FIN_GN) {
        MGDT_CHECK(!quads && out_sumsq && fin.p0 && fin.p1 && fin.o0 && fin.o1 && fin.i0 > 0 && C % fin.i0 == 0,
                   "chan_stats_fin: bad GroupNorm arguments");
    } else {
        MGDT_CHECK(fin.kind == 0, "chan_stats_fin: unknown finaliser %d", fin.kind);
    }
    return chan_stats_impl(x, x_cs, N, H, W, C, quads, out_sum, out_sumsq, ws, ws_bytes, counters, fin, dtype, stream);
}

// ------------------------------------------------------------------ consumer of conv-epilogue statistics
namespace mgdt {
__global__ void __launch_bounds__(CS_THREADS) stats_finish_kernel(double* __restrict__ acc, int R, int N, int H, int W, int C, int Q,
                                                                  int sq, float* __restrict__ out_sum,
                                                                  float* __restrict__ out_sumsq, StatsFin fin) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sm[CS_THREADS * 8];
    const int n = blockIdx.x, K = Q + sq;
    const size_t rs = (size_t)N * K * C;
    double* an = acc + (size_t)n * K * C;
    const bool derive_tot = Q == 5 && ((H | W) & 1) == 0;   // the conv skipped the total plane: the windows partition the image
    auto rd = [&](int i) -> double {   // copies summed in copy order, then reset for the next convolution
        double v = 0.0;
        for (int r0 = 0; r0 < R; r0 += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = r0 + u < R ? __ldcg(&an[(size_t)(r0 + u) * rs + i]) : 0.0;   // loads first
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                v += t[u];
                if (r0 + u < R) an[(size_t)(r0 + u) * rs + i] = 0.0;
            }
        }
        return v;
    };
    for (int i = threadIdx.x; i < K * C; i += CS_THREADS) {
        if (derive_tot && i < C) continue;
        const double v = rd(i);
        if (i < Q * C) out_sum[(size_t)n * Q * C + i] = (float)v;
        else out_sumsq[(size_t)n * C + (i - Q * C)] = (float)v;
    }
    if (derive_tot) {
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += CS_THREADS) {
            const float* q = out_sum + (size_t)n * Q * C + C + c;
            out_sum[(size_t)n * Q * C + c] = (q[0] + q[C]) + (q[2 * C] + q[3 * C]);   // as mgdt_chan_stats does
        }
    }
    if (!fin.kind) return;
    __syncthreads();   // the bodies read this block's own global writes
    if (fin.kind == 1)
        mspa_gate_body(n, out_sum + (size_t)n * Q * C, H, W, C, fin.i0, fin.i1, fin.p0, fin.p1, fin.p2, fin.p3, fin.i2, fin.o0, sm);
    else if (fin.kind == 2)
        grn_scale_body(n, out_sumsq + (size_t)n * C, fin.p0, C, fin.o0, sm);
    else if (fin.kind == 3)
        gn_affine_body(n, out_sum + (size_t)n * Q * C, out_sumsq + (size_t)n * C, C, fin.i0, H * W, fin.f0, fin.p0, fin.p1,
                       fin.o0, fin.o1);
}
}  // namespace mgdt

extern "C" int mgdt_stats_finish(void* acc, int copies, int N, int H, int W, int C, int q, int sq, float* out_sum, float* out_sumsq,
                                 const mgdt_stats_fin* f, void* stream) {
    MGDT_CHECK(acc && copies > 0 && N > 0 && H > 0 && W > 0 && C > 0, "stats_finish: bad arguments");
    MGDT_CHECK((q == 0 || q == 1 || q == 5) && (sq == 0 || sq == 1) && q + sq > 0, "stats_finish: bad plane selection q=%d sq=%d", q, sq);
    MGDT_CHECK((q == 0 || out_sum) && (sq == 0 || out_sumsq), "stats_finish: missing output buffer");
    StatsFin fin{};
    if (f) {
        fin.kind = f->kind; fin.p0 = f->p0; fin.p1 = f->p1; fin.p2 = f->p2; fin.p3 = f->p3;
        fin.i0 = f->i0; fin.i1 = f->i1; fin.i2 = f->i2; fin.f0 = f->f0; fin.o0 = f->o0; fin.o1 = f->o1;
    }
    if (fin.kind == MGDT_FIN_GATE) {
        MGDT_CHECK(q == 5 && fin.p0 && fin.p1 && fin.p2 && fin.p3 && fin.o0 && fin.i0 > 0 && C % fin.i0 == 0 && fin.i2 > 0,
                   "stats_finish: bad SPR gate arguments");
        const int ow = C / fin.i0;
        MGDT_CHECK(fin.i0 * 5 * ow + fin.i0 * fin.i2 + fin.i0 * ow <= CS_THREADS * 8, "stats_finish: gate scratch does not fit (C=%d)", C);
    } else if (fin.kind == MGDT_FIN_GRN) {
        MGDT_CHECK(sq && fin.p0 && fin.o0, "stats_finish: bad GRN arguments");
    } else if (fin.kind == MGDT_FIN_GN) {
        MGDT_CHECK(q == 1 && sq && fin.p0 && fin.p1 && fin.o0 && fin.o1 && fin.i0 > 0 && C % fin.i0 == 0,
                   "stats_finish: bad GroupNorm arguments");
    } else {
        MGDT_CHECK(fin.kind == 0, "stats_finish: unknown finaliser %d", fin.kind);
    }
    launch_k(stats_finish_kernel, dim3(N), dim3(CS_THREADS), 0, (cudaStream_t)stream, (double*)acc, copies, N, H, W, C, q, sq, out_sum, out_sumsq, fin);
    MGDT_LAUNCH_CHECK("stats_finish");
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
    extern __shared__ float gate_sm[];
    mspa_gate_body(blockIdx.x, stats + (long long)blockIdx.x * 5 * C, H, W, C, G, softmax, w1, b1, w2, b2, hidden, scale, gate_sm);
}

__global__ void grn_scale_kernel(const float* __restrict__ sumsq, const float* __restrict__ gamma, int C,
                                 float* __restrict__ scale) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    grn_scale_body(blockIdx.x, sumsq + (long long)blockIdx.x * C, gamma, C, scale, red);
}

__global__ void gn_affine_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, int C, int groups,
                                 int hw, float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ a, float* __restrict__ b) {
    pdl_trigger();
    pdl_wait();
    gn_affine_body(blockIdx.x, sum + (long long)blockIdx.x * C, sumsq + (long long)blockIdx.x * C, C, groups, hw, eps, gamma,
                   beta, a, b);
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
