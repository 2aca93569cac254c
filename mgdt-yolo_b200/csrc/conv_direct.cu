// Generic CUDA-core implicit-GEMM convolution (fp32 accumulate), NHWC, any k/stride/pad,
// any Cin/Cout, with the fused input transforms and epilogue of mgdt_conv2d.
//
// Role: (1) the fp32 validation mode of the whole path (BASELINE.json north_star: "1e-4 in an
// fp32 validation mode"); (2) the bf16 path for shapes the tcgen05 kernel does not take
// (Cin = 3 stem, Cout in {1, 2, 27, ...}).  The tensor-core path lives in conv_umma2.cu.
//
// Tiling: one CTA = BM output pixels x BN output channels, 256 threads, each thread a 4x4
// register tile; K = taps x Cin is walked in chunks of BK=16 staged through shared memory.
#include "common.cuh"

namespace mgdt {

struct ConvP {
    const void* x; const void* w; const float* bias; void* y;
    const void* pre_add; const float* in_scale; const void* pix_scale; const void* residual;
    int N, H, W, Cin, Cout, Ho, Wo, kh, kw, stride, pad;
    int x_cs, y_cs, add_cs, ps_cs, res_cs, act, in_relu;
    long long M;  // N*Ho*Wo
};

constexpr int BK = 16;
constexpr int NTHREADS = 256;

template <typename T, int BN>
__global__ void __launch_bounds__(NTHREADS) conv_direct_kernel(ConvP p) {
    pdl_trigger();
    pdl_wait();
    constexpr int NT_N = BN / 4;          // threads along N
    constexpr int NT_M = NTHREADS / NT_N; // threads along M
    constexpr int BM = NT_M * 4;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % NT_N, ty = tid / NT_N;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const T* __restrict__ x = (const T*)p.x;
    const T* __restrict__ w = (const T*)p.w;
    const T* __restrict__ padd = (const T*)p.pre_add;
    const T* __restrict__ pps = (const T*)p.pix_scale;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int taps = p.kh * p.kw;
    const int HoWo = p.Ho * p.Wo;

    // Each thread stages A elements for fixed pixel slots: e = tid + r*256 -> m = e % BM, k = e / BM.
    // BM is a multiple of 256 or divides it, so a thread's m-slot(s) are loop invariant.
    constexpr int A_ITERS = BM * BK / NTHREADS;
    constexpr int B_ITERS = (BN * BK + NTHREADS - 1) / NTHREADS;

    for (int t = 0; t < taps; ++t) {
        const int dy = t / p.kw, dx = t % p.kw;
        for (int c0 = 0; c0 < p.Cin; c0 += BK) {
            // ---- stage A[k][m]
#pragma unroll 4
            for (int r = 0; r < A_ITERS; ++r) {
                const int e = tid + r * NTHREADS;
                const int m = e % BM, k = e / BM;
                const long long gm = m0 + m;
                const int c = c0 + k;
                float v = 0.f;
                if (gm < p.M && c < p.Cin) {
                    const int n = (int)(gm / HoWo);
                    const int rem = (int)(gm - (long long)n * HoWo);
                    const int ho = rem / p.Wo, wo = rem - ho * p.Wo;
                    const int hi = ho * p.stride - p.pad + dy, wi = wo * p.stride - p.pad + dx;
                    if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) {
                        const long long pix = ((long long)n * p.H + hi) * p.W + wi;
                        v = ldf(x + pix * p.x_cs + c);
                        if (padd) v += ldf(padd + pix * p.add_cs + c);
                        if (p.in_scale) v *= p.in_scale[(long long)n * p.Cin + c];
                        if (pps) v *= ldf(pps + pix * p.ps_cs);
                        if (p.in_relu) v = fmaxf(v, 0.f);
                    }
                }
                As[k][m] = v;
            }
            // ---- stage B[k][n] from OHWI weights
#pragma unroll
            for (int r = 0; r < B_ITERS; ++r) {
                const int e = tid + r * NTHREADS;
                if (e < BN * BK) {
                    const int k = e % BK, nn = e / BK;
                    const int c = c0 + k, co = n0 + nn;
                    float v = 0.f;
                    if (c < p.Cin && co < p.Cout) v = ldf(w + ((long long)co * taps + t) * p.Cin + c);
                    Bs[k][nn] = v;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue
    T* __restrict__ y = (T*)p.y;
    const T* __restrict__ res = (const T*)p.residual;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gm = m0 + ty * 4 + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = n0 + tx * 4 + j;
            if (co >= p.Cout) continue;
            float v = acc[i][j] + (p.bias ? p.bias[co] : 0.f);
            v = apply_act(v, p.act);
            if (res) v += ldf(res + gm * p.res_cs + co);
            stf(y + gm * p.y_cs + co, v);
        }
    }
}

template <typename T, int BN>
static int launch(const ConvP& p, cudaStream_t s) {
    constexpr int BM = (NTHREADS / (BN / 4)) * 4;
    dim3 grid(cdiv(p.M, BM), cdiv(p.Cout, BN));
    launch_k(conv_direct_kernel<T, BN>, dim3(grid), dim3(NTHREADS), 0, s, p);
    MGDT_LAUNCH_CHECK("conv_direct");
    return 0;
}

int conv2d_direct(const mgdt_conv_args* a, cudaStream_t s) {
    ConvP p;
    p.x = a->x; p.w = a->w; p.bias = a->bias; p.y = a->y;
    p.pre_add = a->pre_add; p.in_scale = a->in_scale; p.pix_scale = a->pix_scale; p.residual = a->residual;
    p.N = a->N; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.Cout = a->Cout;
    p.kh = a->kh; p.kw = a->kw; p.stride = a->stride; p.pad = a->pad;
    p.Ho = (a->H + 2 * a->pad - a->kh) / a->stride + 1;
    p.Wo = (a->W + 2 * a->pad - a->kw) / a->stride + 1;
    p.x_cs = a->x_cs; p.y_cs = a->y_cs; p.add_cs = a->add_cs; p.ps_cs = a->ps_cs; p.res_cs = a->res_cs;
    p.act = a->act; p.in_relu = a->in_relu;
    p.M = (long long)p.N * p.Ho * p.Wo;
    if (p.M == 0) return 0;
    MGDT_DTYPE_SWITCH(a->dtype, T, {
        if (p.Cout <= 8) return launch<T, 8>(p, s);
        if (p.Cout <= 16) return launch<T, 16>(p, s);
        if (p.Cout <= 32) return launch<T, 32>(p, s);
        return launch<T, 64>(p, s);
    });
    return 0;
}

}  // namespace mgdt
