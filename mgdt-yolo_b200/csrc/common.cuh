// Shared device/host helpers for libmgdt_b200.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mgdt_b200.h"

namespace mgdt {

// ---- error plumbing (thread-local message, errno-style negative return codes)
int set_error(int code, const char* fmt, ...);

#define MGDT_CHECK(cond, ...)                                   \
    do {                                                        \
        if (!(cond)) return ::mgdt::set_error(-EINVAL, __VA_ARGS__); \
    } while (0)

extern unsigned long long g_launches;  // kernels enqueued by this library (see mgdt_launch_count)

#define MGDT_LAUNCH_CHECK(name)                                                              \
    do {                                                                                     \
        __atomic_fetch_add(&::mgdt::g_launches, 1ULL, __ATOMIC_RELAXED);                     \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess) return ::mgdt::set_error(-EIO, "%s: %s", name, cudaGetErrorString(e_)); \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch (PDL): every kernel of this library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its CTAs may be scheduled while the previous kernel of the
// stream is still draining.  Contract: a kernel calls pdl_trigger() first (lets ITS successor be scheduled early)
// and pdl_wait() before it reads or writes any global memory another kernel may touch; pdl_wait() returns once every
// prerequisite grid has completed and its writes are visible.  MGDT_PDL=0 in the environment launches plainly.
extern int g_pdl;  // -1 unread, 0 off, 1 on
int pdl_enabled();
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- storage-type helpers: T is float or __nv_bfloat16, arithmetic is always fp32
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case MGDT_ACT_SILU: return v / (1.0f + expf(-v));
        case MGDT_ACT_RELU: return fmaxf(v, 0.0f);
        case MGDT_ACT_SIGMOID: return sigmoidf_(v);
        case MGDT_ACT_HSIGMOID: return fminf(fmaxf(v + 3.0f, 0.0f), 6.0f) / 6.0f;
        case MGDT_ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        default: return v;
    }
}

#define MGDT_ERF_A 1.1281433796402367f
#define MGDT_ERF_B 0.1040811854143687f
#define MGDT_ERF_C (-0.0017864744413891597f)

// Epilogue activations on the SFU (bf16 outputs: 2^-9 relative rounding dominates their error).
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <int ACT> __device__ __forceinline__ float act_fast(float v) {
    if (ACT == MGDT_ACT_SILU) return v * fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);         // x * sigmoid(x)
    if (ACT == MGDT_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == MGDT_ACT_SIGMOID) return fmaf(0.5f, tanh_fast(0.5f * v), 0.5f);
    if (ACT == MGDT_ACT_HSIGMOID) return __saturatef(fmaf(v, 1.0f / 6.0f, 0.5f));
    if (ACT == MGDT_ACT_GELU) {
        // erf-GELU with erf(z) ~= tanh(z * (a + b z^2 + c z^4)), |z| clamped to 5 (minimax fit: |erf error| < 3.7e-5,
        // |gelu error| < 5.5e-5, below tanh.approx's own 2^-11 and far below bf16 output rounding): one SFU op
        const float z = fminf(fmaxf(v * 0.70710678118654752440f, -5.0f), 5.0f);
        const float u = z * z;
        const float t = tanh_fast(z * fmaf(fmaf(MGDT_ERF_C, u, MGDT_ERF_B), u, MGDT_ERF_A));
        const float h = 0.5f * v;
        return fmaf(h, t, h);
    }
    return v;
}

__device__ __forceinline__ float act_fast_rt(float v, int act) {
    switch (act) {
        case MGDT_ACT_SILU: return act_fast<MGDT_ACT_SILU>(v);
        case MGDT_ACT_RELU: return act_fast<MGDT_ACT_RELU>(v);
        case MGDT_ACT_SIGMOID: return act_fast<MGDT_ACT_SIGMOID>(v);
        case MGDT_ACT_HSIGMOID: return act_fast<MGDT_ACT_HSIGMOID>(v);
        case MGDT_ACT_GELU: return act_fast<MGDT_ACT_GELU>(v);
        default: return v;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Dispatch a templated launch on the runtime dtype.
#define MGDT_DTYPE_SWITCH(dtype, T, ...)                                      \
    do {                                                                      \
        if ((dtype) == MGDT_F32) {                                            \
            using T = float;                                                  \
            __VA_ARGS__;                                                      \
        } else if ((dtype) == MGDT_BF16) {                                    \
            using T = __nv_bfloat16;                                          \
            __VA_ARGS__;                                                      \
        } else {                                                              \
            return ::mgdt::set_error(-EINVAL, "unsupported dtype %d", (int)(dtype)); \
        }                                                                     \
    } while (0)

}  // namespace mgdt
