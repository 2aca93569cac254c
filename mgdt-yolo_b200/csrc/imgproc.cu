// Pre- and post-processing rows of the hot path, bit-exact with the reference's CPU code.
//
//   mgdt_letterbox_u8   LetterBox (yolo/data/augment.py:538-593: cv2.resize INTER_LINEAR + constant border 114)
//                       fused with the BGR->RGB / HWC->CHW of BasePredictor.preprocess (predictor.py:121-125).
//                       cv2's uint8 bilinear is fixed point: 11-bit coefficients cvRound(w * 2048), horizontal pass in
//                       int, vertical pass ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2; the coefficients come
//                       from float((d + 0.5) * scale - 0.5) evaluated in double -- reproduced here operation by
//                       operation (this file is compiled with -fmad=false so no multiply-add is contracted).
//   mgdt_scale_boxes    ops.scale_boxes + clip_boxes (yolo/utils/ops.py:90-117, 269-285) on the packed NMS output.
#include "common.cuh"

namespace mgdt {

struct LbP {
    const uint8_t* src; int h0, w0, pitch;
    uint8_t* dst; int H, W, new_h, new_w, top, left, swap_rb, pad_value, out_hwc;
    double scale_x, scale_y;
};

__device__ __forceinline__ int sat_short(float v) {
    const int r = __float2int_rn(v);   // cvRound: round half to even
    return r < -32768 ? -32768 : (r > 32767 ? 32767 : r);
}

// source index + the two 11-bit weights of destination coordinate d (resize.cpp, resizeGeneric_ set-up loops)
__device__ __forceinline__ void lin_coef(int d, double scale, int n_src, bool clamp_f, int& s, int& c0, int& c1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (clamp_f) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    }
    c0 = sat_short((1.f - f) * 2048.f);
    c1 = sat_short(f * 2048.f);
}

__global__ void __launch_bounds__(256) letterbox_u8_kernel(LbP p) {
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= p.W) return;
    int v[3] = {p.pad_value, p.pad_value, p.pad_value};
    const int dx = x - p.left, dy = y - p.top;
    if (dx >= 0 && dx < p.new_w && dy >= 0 && dy < p.new_h) {
        int sx, a0, a1, sy, b0, b1;
        lin_coef(dx, p.scale_x, p.w0, true, sx, a0, a1);
        lin_coef(dy, p.scale_y, p.h0, false, sy, b0, b1);
        const int sx1 = min(sx + 1, p.w0 - 1);
        const int y0 = min(max(sy, 0), p.h0 - 1), y1 = min(max(sy + 1, 0), p.h0 - 1);
        const uint8_t* r0 = p.src + (size_t)y0 * p.pitch;
        const uint8_t* r1 = p.src + (size_t)y1 * p.pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = (int)r0[sx * 3 + c] * a0 + (int)r0[sx1 * 3 + c] * a1;
            const int h1 = (int)r1[sx * 3 + c] * a0 + (int)r1[sx1 * 3 + c] * a1;
            const int o = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v[c] = min(max(o, 0), 255);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int co = p.swap_rb ? 2 - c : c;
        if (p.out_hwc) p.dst[((size_t)y * p.W + x) * 3 + co] = (uint8_t)v[c];
        else p.dst[((size_t)co * p.H + y) * p.W + x] = (uint8_t)v[c];
    }
}

__global__ void __launch_bounds__(256) scale_boxes_kernel(float* __restrict__ dets, int row, const int32_t* __restrict__ counts,
                                                          int max_rows, const float* __restrict__ prm, int N) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    const int cnt = counts ? counts[n] : max_rows;
    if (n >= N || i >= cnt || i >= max_rows) return;
    const float gain = prm[n * 5 + 0], padw = prm[n * 5 + 1], padh = prm[n * 5 + 2], h0 = prm[n * 5 + 3], w0 = prm[n * 5 + 4];
    float* b = dets + ((size_t)n * max_rows + i) * row;
    // boxes[..., [0, 2]] -= pad[0]; boxes[..., [1, 3]] -= pad[1]; boxes[..., :4] /= gain; clip (ops.py:111-116)
    b[0] = fminf(fmaxf((b[0] - padw) / gain, 0.f), w0);
    b[1] = fminf(fmaxf((b[1] - padh) / gain, 0.f), h0);
    b[2] = fminf(fmaxf((b[2] - padw) / gain, 0.f), w0);
    b[3] = fminf(fmaxf((b[3] - padh) / gain, 0.f), h0);
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_letterbox_u8(const void* src, int h0, int w0, int pitch, void* dst, int H, int W, int new_h, int new_w,
                                 int top, int left, int swap_rb, int pad_value, int out_hwc, void* stream) {
    MGDT_CHECK(src && dst, "letterbox: null pointer");
    MGDT_CHECK(h0 > 0 && w0 > 0 && pitch >= 3 * w0 && H > 0 && W > 0 && new_h > 0 && new_w > 0, "letterbox: bad shape");
    MGDT_CHECK(top >= 0 && left >= 0 && top + new_h <= H && left + new_w <= W, "letterbox: resized image does not fit the canvas");
    MGDT_CHECK(pad_value >= 0 && pad_value <= 255, "letterbox: bad pad value");
    LbP p;
    p.src = (const uint8_t*)src; p.h0 = h0; p.w0 = w0; p.pitch = pitch; p.dst = (uint8_t*)dst; p.H = H; p.W = W;
    p.new_h = new_h; p.new_w = new_w; p.top = top; p.left = left; p.swap_rb = swap_rb; p.pad_value = pad_value;
    p.out_hwc = out_hwc;
    // cv::resize: inv_scale = dsize / ssize (double), scale = 1. / inv_scale
    p.scale_x = 1.0 / ((double)new_w / (double)w0);
    p.scale_y = 1.0 / ((double)new_h / (double)h0);
    launch_k(letterbox_u8_kernel, dim3(cdiv(W, 256), H), dim3(256), 0, (cudaStream_t)stream, p);
    MGDT_LAUNCH_CHECK("letterbox_u8");
    return 0;
}

extern "C" int mgdt_scale_boxes(float* dets, int row_stride, const int32_t* counts, int N, int max_rows, const float* params,
                                void* stream) {
    MGDT_CHECK(dets && params, "scale_boxes: null pointer");
    MGDT_CHECK(N > 0 && max_rows >= 0 && row_stride >= 4, "scale_boxes: bad shape");
    if (max_rows == 0) return 0;
    launch_k(scale_boxes_kernel, dim3(cdiv(max_rows, 256), N), dim3(256), 0, (cudaStream_t)stream, dets, row_stride, counts,
             max_rows, params, N);
    MGDT_LAUNCH_CHECK("scale_boxes");
    return 0;
}
