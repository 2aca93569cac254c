// Post-NMS result path: Boxes views and the validator's detection <-> label matching.
// Compiled with -fmad=false: every box operation is a single IEEE fp32 operation in the reference's order, so IoU
// threshold decisions and converted coordinates are bit-exact with the torch CPU reference.
#include "common.cuh"

#include <limits.h>

namespace mgdt {

// ------------------------------------------------------------------ Boxes.xywh / xyxyn / xywhn
// mode bit 0: xyxy -> xywh (ops.xyxy2xywh, yolo/utils/ops.py:355-358); bit 1: divide x by w, y by h (results.py:418-430)
__global__ void box_convert_kernel(const float* __restrict__ boxes, int row_stride, int n, int mode, float w, float h,
                                   float* __restrict__ out) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* b = boxes + (size_t)i * row_stride;
    float v0 = b[0], v1 = b[1], v2 = b[2], v3 = b[3];
    if (mode & 1) {
        const float cx = (v0 + v2) / 2.0f, cy = (v1 + v3) / 2.0f, bw = v2 - v0, bh = v3 - v1;
        v0 = cx; v1 = cy; v2 = bw; v3 = bh;
    }
    if (mode & 2) { v0 = v0 / w; v2 = v2 / w; v1 = v1 / h; v3 = v3 / h; }
    float* o = out + (size_t)i * 4;
    o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
}

// ------------------------------------------------------------------ DetectionValidator._process_batch
// One block per image.  Phase 1: every detection finds its highest-IoU class-matching label (the first row np.unique
// keeps of the IoU-sorted match list, val.py:169-170; the same label at every IoU level, valid at level i iff that IoU
// >= iouv[i]).  Phase 2: every label keeps, per level, the LOWEST-index detection that chose it (np.unique over the
// label column of the detection-ordered list, val.py:173 -- the reference's second IoU sort is commented out).
// Phase 3: those detections are `correct`.  IoU as metrics.box_iou (metrics.py:67-72) computes it.
__global__ void __launch_bounds__(256) match_batch_kernel(const float* __restrict__ dets, int det_stride,
                                                          const int* __restrict__ det_counts, int max_det,
                                                          const float* __restrict__ labels, const int* __restrict__ lab_counts,
                                                          int max_lab, const float* __restrict__ iouv, int niou,
                                                          unsigned char* __restrict__ correct) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ int mb_sm[];
    const int n = blockIdx.x;
    const int nd = det_counts ? min(det_counts[n], max_det) : max_det;
    const int nl = lab_counts ? min(lab_counts[n], max_lab) : max_lab;
    int* win = mb_sm;                                   // [max_lab][niou] lowest detection index per (label, level)
    int* best_l = win + max_lab * niou;                 // [max_det]
    float* best_iou = reinterpret_cast<float*>(best_l + max_det);
    const float* dn = dets + (size_t)n * max_det * det_stride;
    const float* ln = labels + (size_t)n * max_lab * 5;
    unsigned char* cn = correct + (size_t)n * max_det * niou;
    for (int i = threadIdx.x; i < nl * niou; i += blockDim.x) win[i] = INT_MAX;
    for (int d = threadIdx.x; d < nd; d += blockDim.x) {
        const float bx1 = dn[d * det_stride], by1 = dn[d * det_stride + 1], bx2 = dn[d * det_stride + 2],
                    by2 = dn[d * det_stride + 3], bc = dn[d * det_stride + 5];
        const float area2 = (bx2 - bx1) * (by2 - by1);
        float bi = -1.0f;
        int bl = -1;
        for (int l = 0; l < nl; ++l) {
            if (ln[l * 5] != bc) continue;
            const float ax1 = ln[l * 5 + 1], ay1 = ln[l * 5 + 2], ax2 = ln[l * 5 + 3], ay2 = ln[l * 5 + 4];
            const float iw = fmaxf(fminf(ax2, bx2) - fmaxf(ax1, bx1), 0.0f);
            const float ih = fmaxf(fminf(ay2, by2) - fmaxf(ay1, by1), 0.0f);
            const float inter = iw * ih;
            const float area1 = (ax2 - ax1) * (ay2 - ay1);
            const float iou = inter / (area1 + area2 - inter + 1e-7f);
            if (iou > bi) { bi = iou; bl = l; }
        }
        best_l[d] = bl;
        best_iou[d] = bi;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nd; d += blockDim.x) {
        const int bl = best_l[d];
        if (bl < 0) continue;
        for (int i = 0; i < niou; ++i)
            if (best_iou[d] >= iouv[i]) atomicMin(&win[bl * niou + i], d);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < max_det * niou; e += blockDim.x) {
        const int d = e / niou, i = e - d * niou;
        unsigned char c = 0;
        if (d < nd) {
            const int bl = best_l[d];
            c = (bl >= 0 && best_iou[d] >= iouv[i] && win[bl * niou + i] == d) ? 1 : 0;
        }
        cn[e] = c;
    }
}

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_box_convert(const float* boxes, int row_stride, int n, int mode, float w, float h, float* out, void* stream) {
    MGDT_CHECK(n >= 0 && row_stride >= 4 && mode >= 0 && mode <= 3, "box_convert: bad arguments");
    if (n == 0) return 0;
    MGDT_CHECK(boxes && out, "box_convert: null pointer");
    MGDT_CHECK(!(mode & 2) || (w > 0.f && h > 0.f), "box_convert: normalisation needs a positive image size");
    launch_k(box_convert_kernel, dim3(cdiv(n, 128)), dim3(128), 0, (cudaStream_t)stream, boxes, row_stride, n, mode, w, h, out);
    MGDT_LAUNCH_CHECK("box_convert");
    return 0;
}

extern "C" int mgdt_match_batch(const float* dets, int det_stride, const int32_t* det_counts, int max_det, const float* labels,
                                const int32_t* lab_counts, int max_lab, const float* iouv, int niou, uint8_t* correct, int N,
                                void* stream) {
    MGDT_CHECK(N >= 0 && max_det >= 0 && max_lab >= 0 && niou > 0 && niou <= 32 && det_stride >= 6, "match_batch: bad arguments");
    if (N == 0 || max_det == 0) return 0;
    MGDT_CHECK(dets && iouv && correct && (labels || max_lab == 0), "match_batch: null pointer");
    const size_t smem = sizeof(int) * ((size_t)max_lab * niou + 2 * (size_t)max_det);
    MGDT_CHECK(smem <= 200 * 1024, "match_batch: %d labels x %d levels + %d detections exceed the shared-memory scratch", max_lab, niou, max_det);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(match_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-EIO, "match_batch: smem attr: %s", cudaGetErrorString(e));
    }
    launch_k(match_batch_kernel, dim3(N), dim3(256), smem, (cudaStream_t)stream, dets, det_stride, (const int*)det_counts, max_det, labels,
             (const int*)lab_counts, max_lab, iouv, niou, (unsigned char*)correct);
    MGDT_LAUNCH_CHECK("match_batch");
    return 0;
}
