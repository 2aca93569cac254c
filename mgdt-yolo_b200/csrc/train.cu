// Row f3: the detection training criterion and the optimizer-side element-wise work.
//
//   mgdt_v8_loss      v8DetectionLoss.__call__ (yolo/utils/loss.py:159-208) on the concatenated head output (B, no, A):
//                     DFL-expectation decode (loss.py:150-157), HeuristicPositiveSampleAssigner_v1 -> TaskAlignedAssigner
//                     (yolo/utils/tal.py:81-142, 171-353: CIoU overlaps, annealed alpha, top-10, select_highest_overlaps on
//                     the alignment metric, normalised target scores), BCE + CIoU + DFL losses (loss.py:60-92,
//                     yolo/utils/metrics.py:75-128) AND their gradient with respect to the head output, in one pass --
//                     the (b, n_gt, anchors) metric tensors of the reference exist only as one fp32 scratch pair.
//   mgdt_ema_update   ModelEMA.update (yolo/utils/torch_utils.py:347-358) over one flat fp32 buffer.
//   mgdt_sgd_step     clip_grad_norm_(10) + SGD(momentum, nesterov) with the three parameter groups of
//                     build_optimizer (yolo/engine/trainer.py:614-650, 462-470) over one flat bucket.
//   mgdt_sumsq        sum of squares of a flat buffer (the gradient norm), fp64 accumulation.
//
// All sums that decide a result are accumulated in fp64 (atomics), so results do not depend on arrival order.
#include "common.cuh"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mgdt {

struct LossP {
    const float *pred, *anchors, *strides, *gt;
    float *align, *ov;            // (B, G, A) scratch
    unsigned char* flags;          // (B, G, A): bit 0 = anchor inside a valid gt box, bit 1 = among the gt's top-k
    int* idx;                      // (B, A) assigned gt or -1
    float *pbox;                   // (B, A, 4) decoded predicted boxes, grid units
    float *pos;                    // (B, G, 2) max alignment metric / max overlap over the gt's positives
    float *tscore;                 // (B, A) normalised target score of the assigned class (0 for background)
    double* acc;                   // [0] target_scores_sum, [1] box, [2] cls, [3] dfl
    float *loss3, *grad;
    float *out_tbox; int* out_label;   // optional assigner outputs (tests): (B, A, 4) pixels, (B, A)
    int B, A, G, nc, R, no, topk;
    float alpha, beta, eps, box_gain, cls_gain, dfl_gain;
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// bbox_iou(box1, box2, xywh=False, CIoU=True) of metrics.py:75-128 (eps 1e-7), same operation order
__device__ __forceinline__ float ciou(float ax1, float ay1, float ax2, float ay2, float bx1, float by1, float bx2, float by2) {
    const float eps = 1e-7f;
    const float w1 = ax2 - ax1, h1 = ay2 - ay1 + eps, w2 = bx2 - bx1, h2 = by2 - by1 + eps;
    const float inter = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1), 0.f) * fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1), 0.f);
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(ax2, bx2) - fminf(ax1, bx1), ch = fmaxf(ay2, by2) - fminf(ay1, by1);
    const float c2 = cw * cw + ch * ch + eps;
    const float sx = bx1 + bx2 - ax1 - ax2, sy = by1 + by2 - ay1 - ay2;
    const float rho2 = (sx * sx + sy * sy) / 4.f;
    const float da = atanf(w2 / h2) - atanf(w1 / h1);
    const float v = 0.4052847345693511f * da * da;   // 4 / pi^2
    const float al = v / (v - iou + (1.f + eps));
    return iou - (rho2 / c2 + v * al);
}

// ---- K0: DFL-expectation decode of the predicted boxes (loss.py:150-157 + dist2bbox), grid units
__global__ void loss_decode_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.A) return;
    const int b = i / p.A, a = i - b * p.A;
    const float* z = p.pred + (size_t)b * p.no * p.A + a;
    float d[4];
    for (int s = 0; s < 4; ++s) {
        float m = -INFINITY;
        for (int k = 0; k < p.R; ++k) m = fmaxf(m, z[(size_t)(s * p.R + k) * p.A]);
        float den = 0.f, num = 0.f;
        for (int k = 0; k < p.R; ++k) {
            const float e = expf(z[(size_t)(s * p.R + k) * p.A] - m);
            den += e;
            num += e * (float)k;
        }
        d[s] = num / den;
    }
    const float ax = p.anchors[2 * a], ay = p.anchors[2 * a + 1];
    float* o = p.pbox + (size_t)i * 4;
    o[0] = ax - d[0]; o[1] = ay - d[1]; o[2] = ax + d[2]; o[3] = ay + d[3];
    if (b == 0 && a == 0) { p.acc[0] = 0.0; p.acc[1] = 0.0; p.acc[2] = 0.0; p.acc[3] = 0.0; }
}

// ---- K1: alignment metric and overlaps of every (gt, anchor) pair (tal.py:240-268)
__global__ void tal_metrics_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    const int a = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y, b = blockIdx.z;
    if (a >= p.A) return;
    const float* t = p.gt + ((size_t)b * p.G + g) * 5;
    const float x1 = t[1], y1 = t[2], x2 = t[3], y2 = t[4];
    const bool valid = (x1 + y1 + x2 + y2) > 0.f;            // mask_gt (loss.py:183)
    const size_t o = ((size_t)b * p.G + g) * p.A + a;
    float al = 0.f, ov = 0.f;
    unsigned char fl = 0;
    if (valid) {
        const float s = p.strides[a], ax = p.anchors[2 * a] * s, ay = p.anchors[2 * a + 1] * s;
        const float dmin = fminf(fminf(ax - x1, ay - y1), fminf(x2 - ax, y2 - ay));   // select_candidates_in_gts
        if (dmin > 1e-9f) {
            fl = 1;
            const float* pb = p.pbox + ((size_t)b * p.A + a) * 4;
            ov = fmaxf(ciou(x1, y1, x2, y2, pb[0] * s, pb[1] * s, pb[2] * s, pb[3] * s), 0.f);
            const int label = (int)t[0];
            const float sc = sigm(p.pred[((size_t)b * p.no + 4 * p.R + label) * p.A + a]);
            al = powf(sc, p.alpha) * powf(ov, p.beta);
        }
    }
    p.align[o] = al; p.ov[o] = ov; p.flags[o] = fl;
}

// ---- K2: top-k anchors of every gt by alignment metric (tal.py:270-305); ties: smaller anchor index first
__global__ void __launch_bounds__(256) tal_topk_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float row[];
    __shared__ float bv[8];
    __shared__ int bi[8];
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const float* t = p.gt + ((size_t)b * p.G + g) * 5;
    if (!((t[1] + t[2] + t[3] + t[4]) > 0.f)) return;     // invalid gt: its top-k indices are masked to anchor 0 and dropped (count > 1)
    const size_t o = ((size_t)b * p.G + g) * p.A;
    for (int a = tid; a < p.A; a += 256) row[a] = p.align[o + a];
    __syncthreads();
    for (int k = 0; k < p.topk && k < p.A; ++k) {
        float v = -1.f;
        int ia = 0x7fffffff;
        for (int a = tid; a < p.A; a += 256) {
            const float x = row[a];
            if (x > v) { v = x; ia = a; }
        }
        for (int off = 16; off; off >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, off);
            const int i2 = __shfl_xor_sync(0xffffffffu, ia, off);
            if (v2 > v || (v2 == v && i2 < ia)) { v = v2; ia = i2; }
        }
        if ((tid & 31) == 0) { bv[tid >> 5] = v; bi[tid >> 5] = ia; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; ++w)
                if (bv[w] > v || (bv[w] == v && bi[w] < ia)) { v = bv[w]; ia = bi[w]; }
            row[ia] = -2.f;                                   // taken
            p.flags[o + ia] |= 2;
        }
        __syncthreads();
    }
}

// ---- K3: one gt per anchor (select_highest_overlaps on the alignment metric, tal.py:28-54, 213)
__global__ void tal_resolve_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.A) return;
    const int b = i / p.A, a = i - b * p.A;
    int cnt = 0, first = -1, best = 0;
    float bestv = -1.f;
    for (int g = 0; g < p.G; ++g) {
        const size_t o = ((size_t)b * p.G + g) * p.A + a;
        if (p.flags[o] == 3) { if (!cnt) first = g; ++cnt; }
        const float v = p.align[o];
        if (v > bestv) { bestv = v; best = g; }               // argmax over ALL gts, first maximum
    }
    p.idx[i] = cnt > 1 ? best : first;
}

// ---- K4: per-gt maxima over its positives (tal.py:219-221)
__global__ void __launch_bounds__(256) tal_posmax_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ float sa[8], so[8];
    const int g = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t o = ((size_t)b * p.G + g) * p.A;
    float ma = 0.f, mo = 0.f;
    for (int a = tid; a < p.A; a += 256)
        if (p.idx[(size_t)b * p.A + a] == g) { ma = fmaxf(ma, p.align[o + a]); mo = fmaxf(mo, p.ov[o + a]); }
    for (int off = 16; off; off >>= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, off));
        mo = fmaxf(mo, __shfl_xor_sync(0xffffffffu, mo, off));
    }
    if ((tid & 31) == 0) { sa[tid >> 5] = ma; so[tid >> 5] = mo; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { ma = fmaxf(ma, sa[w]); mo = fmaxf(mo, so[w]); }
        p.pos[((size_t)b * p.G + g) * 2] = ma;
        p.pos[((size_t)b * p.G + g) * 2 + 1] = mo;
    }
}

// ---- K5: normalised target score per anchor (tal.py:222-223) and target_scores_sum
__global__ void __launch_bounds__(256) tal_targets_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ double red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float ts = 0.f;
    if (i < p.B * p.A) {
        const int b = i / p.A, a = i - b * p.A, g = p.idx[i];
        if (g >= 0) {
            const size_t o = ((size_t)b * p.G + g) * p.A + a;
            const float* ps = p.pos + ((size_t)b * p.G + g) * 2;
            ts = p.align[o] * ps[1] / (ps[0] + p.eps);
        }
        p.tscore[i] = ts;
        if (p.out_label) {
            const float* t = p.gt + ((size_t)b * p.G + max(g, 0)) * 5;
            p.out_label[i] = g >= 0 ? max((int)t[0], 0) : -1;
            for (int j = 0; j < 4; ++j) p.out_tbox[(size_t)i * 4 + j] = g >= 0 ? t[1 + j] : 0.f;
        }
    }
    double d = (double)ts;
    for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) d += red[w];
        if (d != 0.0) atomicAdd(&p.acc[0], d);
    }
}

// ---- K6: the three losses and d(total * B) / d(pred) (loss.py:190-205, 60-92; metrics.py:75-128)
__global__ void __launch_bounds__(256) loss_main_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ double red[3][8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double tss_d = p.acc[0] > 1.0 ? p.acc[0] : 1.0;        // max(target_scores.sum(), 1)
    const float inv = (float)(1.0 / tss_d);
    float l_box = 0.f, l_cls = 0.f, l_dfl = 0.f;
    if (i < p.B * p.A) {
        const int b = i / p.A, a = i - b * p.A, g = p.idx[i];
        const float ts = p.tscore[i];
        const float* z = p.pred + (size_t)b * p.no * p.A + a;
        float* gz = p.grad ? p.grad + (size_t)b * p.no * p.A + a : nullptr;
        const float* t = p.gt + ((size_t)b * p.G + max(g, 0)) * 5;
        const int label = g >= 0 ? max((int)t[0], 0) : -1;
        // classification: BCEWithLogits(pred_scores, target_scores) summed
        const float gs = p.cls_gain * (float)p.B * inv;
        for (int c = 0; c < p.nc; ++c) {
            const float x = z[(size_t)(4 * p.R + c) * p.A];
            const float y = c == label ? ts : 0.f;
            l_cls += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
            if (gz) gz[(size_t)(4 * p.R + c) * p.A] = (sigm(x) - y) * gs;
        }
        if (g >= 0) {
            const float s = p.strides[a], ax = p.anchors[2 * a], ay = p.anchors[2 * a + 1];
            const float w = ts;                                               // target_scores.sum(-1)
            const float tx1 = t[1] / s, ty1 = t[2] / s, tx2 = t[3] / s, ty2 = t[4] / s;   // target_bboxes /= stride_tensor
            // softmax of the four sides and their expectation
            float prob[4][32], d[4];
            for (int sd = 0; sd < 4; ++sd) {
                float m = -INFINITY;
                for (int k = 0; k < p.R; ++k) m = fmaxf(m, z[(size_t)(sd * p.R + k) * p.A]);
                float den = 0.f;
                for (int k = 0; k < p.R; ++k) { prob[sd][k] = expf(z[(size_t)(sd * p.R + k) * p.A] - m); den += prob[sd][k]; }
                float e = 0.f;
                for (int k = 0; k < p.R; ++k) { prob[sd][k] /= den; e += prob[sd][k] * (float)k; }
                d[sd] = e;
            }
            const float px1 = ax - d[0], py1 = ay - d[1], px2 = ax + d[2], py2 = ay + d[3];
            // CIoU(pred, target) and its gradient with respect to the predicted corners (alpha is a constant: no_grad)
            const float eps = 1e-7f;
            const float w1 = px2 - px1, h1 = py2 - py1 + eps, w2 = tx2 - tx1, h2 = ty2 - ty1 + eps;
            const float iw = fminf(px2, tx2) - fmaxf(px1, tx1), ih = fminf(py2, ty2) - fmaxf(py1, ty1);
            const float iwc = fmaxf(iw, 0.f), ihc = fmaxf(ih, 0.f);
            const float inter = iwc * ihc;
            const float uni = w1 * h1 + w2 * h2 - inter + eps;
            const float iou = inter / uni;
            const float cw = fmaxf(px2, tx2) - fminf(px1, tx1), ch = fmaxf(py2, ty2) - fminf(py1, ty1);
            const float c2 = cw * cw + ch * ch + eps;
            const float sx = tx1 + tx2 - px1 - px2, sy = ty1 + ty2 - py1 - py2;
            const float rho2 = (sx * sx + sy * sy) / 4.f;
            const float da = atanf(w2 / h2) - atanf(w1 / h1);
            const float v = 0.4052847345693511f * da * da;
            const float alp = v / (v - iou + (1.f + eps));
            const float ci = iou - (rho2 / c2 + v * alp);
            l_box = (1.f - ci) * w;
            if (gz) {
                // derivatives with respect to (x1, y1, x2, y2) of the prediction
                const float diw[4] = {(iw >= 0.f && px1 > tx1) ? -1.f : 0.f, 0.f, (iw >= 0.f && px2 < tx2) ? 1.f : 0.f, 0.f};
                const float dih[4] = {0.f, (ih >= 0.f && py1 > ty1) ? -1.f : 0.f, 0.f, (ih >= 0.f && py2 < ty2) ? 1.f : 0.f};
                const float dw1[4] = {-1.f, 0.f, 1.f, 0.f}, dh1[4] = {0.f, -1.f, 0.f, 1.f};
                const float dcw[4] = {px1 < tx1 ? -1.f : 0.f, 0.f, px2 > tx2 ? 1.f : 0.f, 0.f};
                const float dch[4] = {0.f, py1 < ty1 ? -1.f : 0.f, 0.f, py2 > ty2 ? 1.f : 0.f};
                const float dsx[4] = {-1.f, 0.f, -1.f, 0.f}, dsy[4] = {0.f, -1.f, 0.f, -1.f};
                const float wh2 = w1 * w1 + h1 * h1;
                float dbox[4];
                for (int j = 0; j < 4; ++j) {
                    const float dinter = diw[j] * ihc + iwc * dih[j];
                    const float duni = dw1[j] * h1 + w1 * dh1[j] - dinter;
                    const float diou = (dinter * uni - inter * duni) / (uni * uni);
                    const float dc2 = 2.f * cw * dcw[j] + 2.f * ch * dch[j];
                    const float drho = (2.f * sx * dsx[j] + 2.f * sy * dsy[j]) / 4.f;
                    const float dterm = (drho * c2 - rho2 * dc2) / (c2 * c2);
                    const float datan1 = (h1 * dw1[j] - w1 * dh1[j]) / wh2;             // d atan(w1 / h1)
                    const float dv = 0.4052847345693511f * 2.f * da * (-datan1);
                    dbox[j] = -(diou - dterm - alp * dv) * w;                          // d (1 - ciou) * w
                }
                // corners -> distances (x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b) -> logits (softmax expectation)
                const float dd[4] = {-dbox[0], -dbox[1], dbox[2], dbox[3]};
                const float gb = p.box_gain * (float)p.B * inv;
                for (int sd = 0; sd < 4; ++sd)
                    for (int k = 0; k < p.R; ++k) gz[(size_t)(sd * p.R + k) * p.A] = gb * dd[sd] * prob[sd][k] * ((float)k - d[sd]);
            }
            // DFL (loss.py:80-90): target distances clamped to [0, reg_max - 1 - 0.01]
            const float tl4[4] = {ax - tx1, ay - ty1, tx2 - ax, ty2 - ay};
            const float gd = p.dfl_gain * (float)p.B * inv * w * 0.25f;
            float ld = 0.f;
            for (int sd = 0; sd < 4; ++sd) {
                const float tv = fminf(fmaxf(tl4[sd], 0.f), (float)(p.R - 1) - 0.01f);
                const int tl = (int)tv, tr = tl + 1;
                const float wl = (float)tr - tv, wr = 1.f - wl;
                ld += -logf(prob[sd][tl]) * wl - logf(prob[sd][tr]) * wr;
                if (gz)
                    for (int k = 0; k < p.R; ++k)
                        gz[(size_t)(sd * p.R + k) * p.A] += gd * (prob[sd][k] - (k == tl ? wl : 0.f) - (k == tr ? wr : 0.f));
            }
            l_dfl = ld * 0.25f * w;
        } else if (gz) {
            for (int k = 0; k < 4 * p.R; ++k) gz[(size_t)k * p.A] = 0.f;
        }
    }
    double v3[3] = {(double)l_box, (double)l_cls, (double)l_dfl};
    for (int j = 0; j < 3; ++j) {
        double d = v3[j];
        for (int off = 16; off; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        if ((threadIdx.x & 31) == 0) red[j][threadIdx.x >> 5] = d;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double d = 0.0;
        for (int w = 0; w < 8; ++w) d += red[threadIdx.x][w];
        if (d != 0.0) atomicAdd(&p.acc[1 + threadIdx.x], d);
    }
}

__global__ void loss_final_kernel(LossP p) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double tss = p.acc[0] > 1.0 ? p.acc[0] : 1.0;
        p.loss3[0] = (float)(p.acc[1] / tss) * p.box_gain;
        p.loss3[1] = (float)(p.acc[2] / tss) * p.cls_gain;
        p.loss3[2] = (float)(p.acc[3] / tss) * p.dfl_gain;
        p.loss3[3] = (float)p.acc[0];
    }
}

// ------------------------------------------------------------------------------------------- optimizer side
__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ model, size_t n, float d) {
    pdl_trigger();
    pdl_wait();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        ema[i] = ema[i] * d + (1.f - d) * model[i];              // v *= d; v += (1 - d) * msd[k]
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, size_t n, double* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    __shared__ double red[8];
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += (double)x[i] * (double)x[i];
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) s += red[w];
        atomicAdd(out, s);
    }
}

// group[i] in {0: weights with decay, 1: normalisation weights, 2: biases}; lr[3], wd[3]
__global__ void sgd_kernel(float* __restrict__ prm, const float* __restrict__ grad, float* __restrict__ mom, const unsigned char* __restrict__ group,
                           size_t n, float lr0, float lr1, float lr2, float wd0, float wd1, float wd2, float momentum, int nesterov,
                           int first, const double* __restrict__ gnorm_sq, float max_norm, float grad_scale) {
    pdl_trigger();
    pdl_wait();
    float clip = 1.f;
    if (gnorm_sq) {   // clip_grad_norm_: coefficient max_norm / (norm + 1e-6), clamped to 1; the norm is of the (unscaled) gradients
        const float nrm = (float)sqrt(*gnorm_sq) * grad_scale;
        clip = fminf(max_norm / (nrm + 1e-6f), 1.f);
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int gi = group[i];
        const float lr = gi == 0 ? lr0 : gi == 1 ? lr1 : lr2, wd = gi == 0 ? wd0 : gi == 1 ? wd1 : wd2;
        float g = grad[i] * grad_scale * clip;
        const float w = prm[i];
        g += wd * w;
        const float b = first ? g : momentum * mom[i] + g;
        mom[i] = b;
        g = nesterov ? g + momentum * b : b;
        prm[i] = w - lr * g;
    }
}

}  // namespace mgdt

using namespace mgdt;

extern "C" size_t mgdt_v8_loss_ws_bytes(int B, int A, int G) {
    const size_t bga = (size_t)B * (G > 0 ? G : 1) * A, ba = (size_t)B * A;
    size_t n = 0;
    n += bga * 4 * 2;                       // align, ov
    n += (bga + 15) / 16 * 16;              // flags
    n += ba * 4;                            // idx
    n += ba * 16;                           // pbox
    n += (size_t)B * (G > 0 ? G : 1) * 8;   // pos
    n += ba * 4;                            // tscore
    n += 64;                                // acc
    return n + 256;
}

extern "C" int mgdt_v8_loss(const float* pred, const float* anchors, const float* strides, const float* gt, int B, int A, int G, int nc,
                            int reg_max, float alpha, float beta, int topk, float box_gain, float cls_gain, float dfl_gain, float* loss4,
                            float* grad_pred, float* out_tscore, int* out_idx, float* out_tbox, int* out_label, void* ws, size_t ws_bytes,
                            void* stream) {
    MGDT_CHECK(pred && anchors && strides && loss4 && ws, "v8_loss: null pointer");
    MGDT_CHECK(B > 0 && A > 0 && G >= 0 && nc > 0 && reg_max >= 2 && reg_max <= 32, "v8_loss: bad shape (reg_max 2..32)");
    MGDT_CHECK(G == 0 || gt, "v8_loss: null targets");
    MGDT_CHECK(ws_bytes >= mgdt_v8_loss_ws_bytes(B, A, G), "v8_loss: workspace too small");
    MGDT_CHECK((size_t)A * 4 <= 48 * 1024, "v8_loss: more than 12288 anchors");
    cudaStream_t s = (cudaStream_t)stream;
    LossP p{};
    p.pred = pred; p.anchors = anchors; p.strides = strides; p.gt = gt;
    p.B = B; p.A = A; p.G = G; p.nc = nc; p.R = reg_max; p.no = 4 * reg_max + nc; p.topk = topk;
    p.alpha = alpha; p.beta = beta; p.eps = 1e-9f; p.box_gain = box_gain; p.cls_gain = cls_gain; p.dfl_gain = dfl_gain;
    const size_t Ge = G > 0 ? G : 1, bga = (size_t)B * Ge * A, ba = (size_t)B * A;
    unsigned char* w = (unsigned char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    p.acc = (double*)w; w += 64;
    p.align = (float*)w; w += bga * 4;
    p.ov = (float*)w; w += bga * 4;
    p.pbox = (float*)w; w += ba * 16;
    p.pos = (float*)w; w += (size_t)B * Ge * 8;
    p.idx = out_idx ? out_idx : (int*)w; w += ba * 4;
    p.tscore = out_tscore ? out_tscore : (float*)w; w += ba * 4;
    p.flags = w;
    p.loss3 = loss4; p.grad = grad_pred; p.out_tbox = out_tbox; p.out_label = (out_tbox && out_label) ? out_label : nullptr;
    const int nb = (int)((ba + 255) / 256);
    launch_k(loss_decode_kernel, dim3(nb), dim3(256), 0, s, p);
    MGDT_LAUNCH_CHECK("loss_decode");
    if (G > 0) {
        launch_k(tal_metrics_kernel, dim3((A + 255) / 256, G, B), dim3(256), 0, s, p);
        MGDT_LAUNCH_CHECK("tal_metrics");
        launch_k(tal_topk_kernel, dim3(G, B), dim3(256), (size_t)A * 4, s, p);
        MGDT_LAUNCH_CHECK("tal_topk");
    }
    launch_k(tal_resolve_kernel, dim3(nb), dim3(256), 0, s, p);
    MGDT_LAUNCH_CHECK("tal_resolve");
    if (G > 0) {
        launch_k(tal_posmax_kernel, dim3(G, B), dim3(256), 0, s, p);
        MGDT_LAUNCH_CHECK("tal_posmax");
    }
    launch_k(tal_targets_kernel, dim3(nb), dim3(256), 0, s, p);
    MGDT_LAUNCH_CHECK("tal_targets");
    launch_k(loss_main_kernel, dim3(nb), dim3(256), 0, s, p);
    MGDT_LAUNCH_CHECK("loss_main");
    launch_k(loss_final_kernel, dim3(1), dim3(32), 0, s, p);
    MGDT_LAUNCH_CHECK("loss_final");
    return 0;
}

extern "C" int mgdt_ema_update(float* ema, const float* model, size_t n, float decay, void* stream) {
    MGDT_CHECK(ema && model, "ema_update: null pointer");
    if (n == 0) return 0;
    const int nb = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    launch_k(ema_kernel, dim3(nb), dim3(256), 0, (cudaStream_t)stream, ema, model, n, decay);
    MGDT_LAUNCH_CHECK("ema_update");
    return 0;
}

extern "C" int mgdt_sumsq(const float* x, size_t n, double* out, void* stream) {
    MGDT_CHECK(x && out, "sumsq: null pointer");
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(-EIO, "sumsq: memset: %s", cudaGetErrorString(e));
    if (n == 0) return 0;
    const int nb = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    launch_k(sumsq_kernel, dim3(nb), dim3(256), 0, (cudaStream_t)stream, x, n, out);
    MGDT_LAUNCH_CHECK("sumsq");
    return 0;
}

extern "C" int mgdt_sgd_step(float* prm, const float* grad, float* mom, const unsigned char* group, size_t n, const float* lr3,
                             const float* wd3, float momentum, int nesterov, int first_step, const double* gnorm_sq, float max_norm,
                             float grad_scale, void* stream) {
    MGDT_CHECK(prm && grad && mom && group && lr3 && wd3, "sgd_step: null pointer");
    if (n == 0) return 0;
    const int nb = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    launch_k(sgd_kernel, dim3(nb), dim3(256), 0, (cudaStream_t)stream, prm, grad, mom, group, n, lr3[0], lr3[1], lr3[2], wd3[0], wd3[1],
             wd3[2], momentum, nesterov, first_step, gnorm_sq, max_norm, grad_scale);
    MGDT_LAUNCH_CHECK("sgd_step");
    return 0;
}
