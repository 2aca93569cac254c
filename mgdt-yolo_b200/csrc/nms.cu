// Batched NMS for the whole batch in four launches and no host synchronisation.
// Replaces non_max_suppression (yolo/utils/ops.py:136-266) + torchvision.ops.nms.
//
//   1. nms_count    per (image, 256-anchor chunk): number of candidates (score > conf [, class filter])
//   2. nms_compact  deterministic compaction -> 64-bit keys  (~score_bits << 32 | anchor*nc + class)
//                   ascending key order == score descending, ties by candidate index (stable sort)
//   3. nms_rank     rank-by-counting sort of the keys (all SMs; keys are unique) -> sorted keys,
//                   truncated to max_nms.  O(n^2) per image: used while an image can have at most NMS_DIRECT_CAP
//                   candidates (the predictor setting: single label, conf 0.25).
//   3'. bucketed sort for large candidate sets (the validator setting conf 0.001 + multi_label: A*nc candidates):
//                   nms_compact also histograms the keys over 16384 buckets (the top 16 bits of the inverted score:
//                   sign / exponent / 7 mantissa bits, monotone in the key), nms_bscan turns the histogram into
//                   bucket offsets, nms_scatter groups the keys by bucket, nms_brank ranks every key among the
//                   members of ITS bucket only: O(n * bucket size) instead of O(n^2) -- 12,800 candidates spread
//                   over ~1,300 occupied buckets.  Same sorted output, bit for bit (keys are unique).
//   4. nms_scan     one CTA per image: greedy suppression over the sorted candidates in chunks of
//                   NMS_CHUNK; each chunk is first tested against the boxes kept so far, then
//                   resolved internally with a 64-bit IoU bitmask and a find-first-set walk.
//                   Stops as soon as max_det boxes are kept (ops.py:250 keeps only i[:max_det]).
//
// Bit-exactness: every box operation is a single IEEE fp32 op in the reference's order
// (xywh2xyxy ops.py:372-376; boxes + cls*max_wh ops.py:247-248; torchvision's
// inter / (area_a + area_b - inter) > thr).  This file is compiled with -fmad=false.
#include "common.cuh"

namespace mgdt {

constexpr int NMS_T = 256;
constexpr int NMS_CHUNK = 512;       // candidates per scan step == threads of nms_scan
constexpr int NMS_WORDS = NMS_CHUNK / 32;
constexpr int NMS_BUCKETS = 16384;   // (key >> 48) - 0xC000 for scores in [0, 2): 2 exponent-range bits + 7 mantissa bits
constexpr int NMS_DIRECT_CAP = 8192; // images that can hold more candidates than this take the bucketed sort

struct NmsP {
    const float* pred; int N, nc, A; float conf, iou; int multi, agnostic, max_det, max_nms; float max_wh;
    const int32_t* classes; int n_classes;
    int nchunks; int cap;            // cap = max candidates per image
    int* blockcnt;                   // [N][nchunks]
    int* ncand;                      // [N]
    unsigned long long* keys;        // [N][cap]
    unsigned long long* sorted;      // [N][max_nms_eff]
    int sorted_cap;
    int* hist;                       // bucketed sort: [N][NMS_BUCKETS + 1] counts -> exclusive offsets (+ total)
    int* cursor;                     // [N][NMS_BUCKETS] scatter cursors (zero on entry)
    unsigned long long* grouped;     // [N][cap] keys grouped by bucket
};

__device__ __forceinline__ int key_bucket(unsigned long long key) {
    const int b = (int)(key >> 48) - 0xC000;   // ascending bucket == ascending key == descending score
    return b < 0 ? 0 : b;                      // scores >= 2 share bucket 0 (still monotone)
}

__device__ __forceinline__ bool class_ok(const NmsP& p, int j) {
    if (!p.classes) return true;
    for (int i = 0; i < p.n_classes; ++i)
        if (p.classes[i] == j) return true;
    return false;
}

// candidates of anchor a in image n: returns count; if `emit`, writes keys starting at out.
__device__ __forceinline__ int anchor_candidates(const NmsP& p, int n, int a, unsigned long long* out) {
    const float* sc = p.pred + ((long long)n * (4 + p.nc) + 4) * p.A + a;
    int cnt = 0;
    if (p.multi) {
        // xc = amax > conf is implied by any class > conf (ops.py:194,226)
        for (int j = 0; j < p.nc; ++j) {
            const float s = sc[(long long)j * p.A];
            if (s > p.conf && class_ok(p, j)) {
                if (out) out[cnt] = ((unsigned long long)(~__float_as_uint(s)) << 32) | (unsigned)(a * p.nc + j);
                ++cnt;
            }
        }
    } else {
        // best class only: conf, j = cls.max(1) (first maximal index), keep conf > conf_thres (ops.py:229-230)
        float best = sc[0];
        int bj = 0;
        for (int j = 1; j < p.nc; ++j) {
            const float s = sc[(long long)j * p.A];
            if (s > best) { best = s; bj = j; }
        }
        if (best > p.conf && class_ok(p, bj)) {
            if (out) out[0] = ((unsigned long long)(~__float_as_uint(best)) << 32) | (unsigned)(a * p.nc + bj);
            cnt = 1;
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(NMS_T) nms_count(NmsP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ int wsum[NMS_T / 32];
    const int n = blockIdx.y, a = blockIdx.x * NMS_T + threadIdx.x;
    int c = (a < p.A) ? anchor_candidates(p, n, a, nullptr) : 0;
    c = (int)warp_sum((float)c);  // counts <= 32*80, exact in fp32
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < NMS_T / 32; ++i) t += wsum[i];
        p.blockcnt[n * p.nchunks + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(NMS_T) nms_compact(NmsP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ int wpre[NMS_T / 32];
    __shared__ int base_s;
    const int n = blockIdx.y, a = blockIdx.x * NMS_T + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        int b = 0;
        for (int i = 0; i < (int)blockIdx.x; ++i) b += p.blockcnt[n * p.nchunks + i];
        base_s = b;
        if (blockIdx.x == gridDim.x - 1) p.ncand[n] = b + p.blockcnt[n * p.nchunks + blockIdx.x];
    }
    const int c = (a < p.A) ? anchor_candidates(p, n, a, nullptr) : 0;
    // exclusive scan of c over the block
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wpre[wid] = incl;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < wid; ++i) woff += wpre[i];
    const int off = base_s + woff + incl - c;
    if (c > 0) {
        unsigned long long* dst = p.keys + (long long)n * p.cap + off;
        anchor_candidates(p, n, a, dst);
        if (p.hist)
            for (int j = 0; j < c; ++j) atomicAdd(p.hist + (long long)n * (NMS_BUCKETS + 1) + key_bucket(dst[j]), 1);
    }
}

// bucket counts -> exclusive offsets (in place), one CTA per image; hist[n][NMS_BUCKETS] receives the total
__global__ void __launch_bounds__(1024) nms_bscan(NmsP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ int wsum[32];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int* h = p.hist + (long long)n * (NMS_BUCKETS + 1);
    constexpr int PER = NMS_BUCKETS / 1024;
    int v[PER], s = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) { v[j] = h[tid * PER + j]; s += v[j]; }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    int run = wsum[wid] + incl - s;
#pragma unroll
    for (int j = 0; j < PER; ++j) { h[tid * PER + j] = run; run += v[j]; }
    if (tid == 1023) h[NMS_BUCKETS] = run;
}

__global__ void __launch_bounds__(NMS_T) nms_scatter(NmsP p) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, i = blockIdx.x * NMS_T + threadIdx.x;
    if (i >= p.ncand[n]) return;
    const unsigned long long key = p.keys[(long long)n * p.cap + i];
    const int b = key_bucket(key);
    const int pos = p.hist[(long long)n * (NMS_BUCKETS + 1) + b] + atomicAdd(p.cursor + (long long)n * NMS_BUCKETS + b, 1);
    p.grouped[(long long)n * p.cap + pos] = key;
}

// rank of every key = offset of its bucket + number of smaller keys in the bucket
__global__ void __launch_bounds__(NMS_T) nms_brank(NmsP p) {
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.y, i = blockIdx.x * NMS_T + threadIdx.x;
    if (i >= p.ncand[n]) return;
    const unsigned long long* g = p.grouped + (long long)n * p.cap;
    const unsigned long long mine = g[i];
    const int b = key_bucket(mine);
    const int* h = p.hist + (long long)n * (NMS_BUCKETS + 1);
    const int lo = h[b], hi = h[b + 1];
    int rank = lo;
    for (int j = lo; j < hi; ++j) rank += (g[j] < mine) ? 1 : 0;
    if (rank < p.sorted_cap) p.sorted[(long long)n * p.sorted_cap + rank] = mine;
}

__global__ void __launch_bounds__(NMS_T) nms_rank(NmsP p) {
    pdl_trigger();
    pdl_wait();
    __shared__ unsigned long long tile[1024];
    const int n = blockIdx.y;
    const int nc_ = p.ncand[n];
    const int i = blockIdx.x * NMS_T + threadIdx.x;
    if (blockIdx.x * NMS_T >= nc_) return;
    const unsigned long long* k = p.keys + (long long)n * p.cap;
    const unsigned long long mine = (i < nc_) ? k[i] : 0xffffffffffffffffULL;
    int rank = 0;
    for (int t0 = 0; t0 < nc_; t0 += 1024) {
        const int len = min(1024, nc_ - t0);
        __syncthreads();
        for (int j = threadIdx.x; j < len; j += NMS_T) tile[j] = k[t0 + j];
        __syncthreads();
        for (int j = 0; j < len; ++j) rank += (tile[j] < mine) ? 1 : 0;
    }
    if (i < nc_ && rank < p.sorted_cap) p.sorted[(long long)n * p.sorted_cap + rank] = mine;
}

struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ bool iou_gt(const Box& a, float area_a, const Box& b, float area_b, float thr) {
    const float xx1 = fmaxf(a.x1, b.x1), yy1 = fmaxf(a.y1, b.y1);
    const float xx2 = fminf(a.x2, b.x2), yy2 = fminf(a.y2, b.y2);
    const float w = fmaxf(xx2 - xx1, 0.f), h = fmaxf(yy2 - yy1, 0.f);
    const float inter = w * h;
    if (!(inter > 0.f)) return false;   // 0 / u > thr is false for every thr >= 0 (and NaN compares false): skip the divide
    return inter / (area_a + area_b - inter) > thr;
}

__global__ void __launch_bounds__(NMS_CHUNK) nms_scan(NmsP p, float* __restrict__ out, int32_t* __restrict__ counts) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned char smraw[];
    // kept boxes (offset coords) + areas, chunk boxes, chunk masks
    Box* kbox = (Box*)smraw;                                  // [max_det]
    float* karea = (float*)(kbox + p.max_det);                // [max_det]
    Box* cbox = (Box*)(karea + p.max_det);                    // [NMS_CHUNK]
    float* carea = (float*)(cbox + NMS_CHUNK);                // [NMS_CHUNK]
    unsigned* cmask = (unsigned*)(carea + NMS_CHUNK);         // [NMS_CHUNK][NMS_WORDS]
    unsigned* alive = cmask + NMS_CHUNK * NMS_WORDS;          // [NMS_WORDS]
    __shared__ int nkept_s;
    __shared__ int newk[NMS_CHUNK];
    __shared__ int n_new_s;

    const int n = blockIdx.x, tid = threadIdx.x;
    const int total = min(p.ncand[n], p.sorted_cap);
    const unsigned long long* sk = p.sorted + (long long)n * p.sorted_cap;
    const float* pr = p.pred + (long long)n * (4 + p.nc) * p.A;
    if (tid == 0) nkept_s = 0;
    __syncthreads();

    for (int c0 = 0; c0 < total; c0 += NMS_CHUNK) {
        const int idx = c0 + tid;
        const bool valid = idx < total;
        Box raw = {0.f, 0.f, 0.f, 0.f}, ob = {0.f, 0.f, 0.f, 0.f};
        float score = 0.f, area = 0.f;
        int cls = 0;
        if (valid) {
            const unsigned long long key = sk[idx];
            const unsigned cand = (unsigned)(key & 0xffffffffu);
            score = __uint_as_float(~(unsigned)(key >> 32));
            const int a = (int)(cand / (unsigned)p.nc);
            cls = (int)(cand - (unsigned)a * (unsigned)p.nc);
            const float cx = pr[a], cy = pr[(long long)p.A + a], w = pr[2LL * p.A + a], h = pr[3LL * p.A + a];
            raw.x1 = cx - w / 2.f; raw.y1 = cy - h / 2.f; raw.x2 = cx + w / 2.f; raw.y2 = cy + h / 2.f;
            const float off = p.agnostic ? ((float)cls * 0.f) : ((float)cls * p.max_wh);
            ob.x1 = raw.x1 + off; ob.y1 = raw.y1 + off; ob.x2 = raw.x2 + off; ob.y2 = raw.y2 + off;
            area = (ob.x2 - ob.x1) * (ob.y2 - ob.y1);
        }
        cbox[tid] = ob;
        carea[tid] = area;
        // (a) against boxes kept in earlier chunks
        const int nk = nkept_s;
        bool live = valid;
        for (int k = 0; live && k < nk; ++k) live = !iou_gt(kbox[k], karea[k], ob, area, p.iou);
        const unsigned bal = __ballot_sync(0xffffffffu, live);
        if ((tid & 31) == 0) alive[tid >> 5] = bal;
        __syncthreads();
        // (b) intra-chunk mask: bit j of row tid set iff j > tid is suppressed by tid
        for (int wd = 0; wd < NMS_WORDS; ++wd) {
            unsigned m = 0;
            if (live && wd >= (tid >> 5)) {
                // live candidates after this one only (words past the last valid candidate are empty)
                unsigned bits = alive[wd];
                if (wd == (tid >> 5)) bits &= ~((2u << (tid & 31)) - 1u);
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1u;
                    const int j = wd * 32 + b;
                    if (iou_gt(ob, area, cbox[j], carea[j], p.iou)) m |= (1u << b);
                }
            }
            cmask[tid * NMS_WORDS + wd] = m;
        }
        __syncthreads();
        // (c) greedy walk by warp 0: lane w owns word w of the removed set
        if (tid < 32) {
            unsigned removed = 0;
            const unsigned aw = (tid < NMS_WORDS) ? alive[tid] : 0u;
            int nnew = 0;
            int kept_total = nk;
            for (int wd = 0; wd < NMS_WORDS && kept_total < p.max_det; ++wd) {
                while (kept_total < p.max_det) {
                    const unsigned cur = __shfl_sync(0xffffffffu, aw & ~removed, wd);
                    if (cur == 0u) break;
                    const int b = __ffs(cur) - 1;
                    const int i = wd * 32 + b;
                    if (tid == 0) newk[nnew] = i;
                    ++nnew;
                    ++kept_total;
                    if (tid < NMS_WORDS) removed |= cmask[i * NMS_WORDS + tid];
                    if (tid == wd) removed |= (1u << b);  // consumed
                }
            }
            if (tid == 0) n_new_s = nnew;
        }
        __syncthreads();
        // (d) append kept boxes, write output rows
        const int nnew = n_new_s;
        for (int q = tid; q < nnew; q += NMS_CHUNK) {
            const int i = newk[q];
            kbox[nk + q] = cbox[i];
            karea[nk + q] = carea[i];
        }
        // each thread knows its own raw box/score/cls: find its slot
        for (int q = 0; q < nnew; ++q) {
            if (newk[q] == tid) {
                float* o = out + ((long long)n * p.max_det + nk + q) * 6;
                o[0] = raw.x1; o[1] = raw.y1; o[2] = raw.x2; o[3] = raw.y2; o[4] = score; o[5] = (float)cls;
            }
        }
        __syncthreads();
        if (tid == 0) nkept_s = nk + nnew;
        __syncthreads();
        if (nkept_s >= p.max_det) break;
    }
    if (tid == 0) counts[n] = nkept_s;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct NmsLayout { size_t blockcnt, ncand, keys, sorted, hist, cursor, grouped, total; int nchunks, cap, sorted_cap, bucketed; };

static NmsLayout nms_layout(int N, int nc, int A, int multi, int max_nms) {
    NmsLayout L;
    L.nchunks = cdiv(A, NMS_T);
    L.cap = A * ((multi && nc > 1) ? nc : 1);
    L.sorted_cap = L.cap < max_nms ? L.cap : max_nms;
    size_t off = 0;
    L.blockcnt = off; off = align_up(off + sizeof(int) * (size_t)N * L.nchunks, 256);
    L.ncand = off;    off = align_up(off + sizeof(int) * (size_t)N, 256);
    L.keys = off;     off = align_up(off + sizeof(unsigned long long) * (size_t)N * L.cap, 256);
    L.sorted = off;   off = align_up(off + sizeof(unsigned long long) * (size_t)N * L.sorted_cap, 256);
    L.bucketed = L.cap > NMS_DIRECT_CAP ? 1 : 0;
    L.hist = L.cursor = L.grouped = off;
    if (L.bucketed) {   // hist and cursor are adjacent: one memset clears both
        L.hist = off;    off = align_up(off + sizeof(int) * (size_t)N * (NMS_BUCKETS + 1), 256);
        L.cursor = off;  off = align_up(off + sizeof(int) * (size_t)N * NMS_BUCKETS, 256);
        L.grouped = off; off = align_up(off + sizeof(unsigned long long) * (size_t)N * L.cap, 256);
    }
    L.total = off;
    return L;
}

}  // namespace mgdt

using namespace mgdt;

extern "C" size_t mgdt_nms_ws_bytes(int N, int nc, int A, int multi_label, int max_nms) {
    if (N <= 0 || nc <= 0 || A <= 0 || max_nms <= 0) return 0;
    return nms_layout(N, nc, A, multi_label, max_nms).total;
}

extern "C" int mgdt_nms(const float* pred, int N, int nc, int A, float conf_thres, float iou_thres, int multi_label,
                        int agnostic, int max_det, int max_nms, float max_wh, const int32_t* classes, int n_classes,
                        float* out, int32_t* counts, void* ws, size_t ws_bytes, void* stream) {
    MGDT_CHECK(pred && out && counts && ws, "nms: null pointer");
    MGDT_CHECK(N > 0 && nc > 0 && A > 0, "nms: bad shape N=%d nc=%d A=%d", N, nc, A);
    // the reference asserts both thresholds in [0, 1] (ops.py:181-182)
    MGDT_CHECK(conf_thres >= 0.f && conf_thres <= 1.f, "Invalid Confidence threshold %g, valid values are between 0.0 and 1.0", conf_thres);
    MGDT_CHECK(iou_thres >= 0.f && iou_thres <= 1.f, "Invalid IoU %g, valid values are between 0.0 and 1.0", iou_thres);
    MGDT_CHECK(max_det > 0 && max_det <= 4096 && max_nms > 0, "nms: bad max_det/max_nms");
    MGDT_CHECK((long long)A * nc < (1LL << 31), "nms: A*nc too large");
    const int multi = (multi_label && nc > 1) ? 1 : 0;  // multi_label &= nc > 1 (ops.py:200)
    const NmsLayout L = nms_layout(N, nc, A, multi, max_nms);
    MGDT_CHECK(ws_bytes >= L.total, "nms: workspace too small (%zu < %zu)", ws_bytes, L.total);
    NmsP p;
    p.pred = pred; p.N = N; p.nc = nc; p.A = A; p.conf = conf_thres; p.iou = iou_thres; p.multi = multi;
    p.agnostic = agnostic; p.max_det = max_det; p.max_nms = max_nms; p.max_wh = max_wh;
    p.classes = (classes && n_classes > 0) ? classes : nullptr; p.n_classes = n_classes;
    p.nchunks = L.nchunks; p.cap = L.cap; p.sorted_cap = L.sorted_cap;
    unsigned char* base = (unsigned char*)ws;
    p.blockcnt = (int*)(base + L.blockcnt); p.ncand = (int*)(base + L.ncand);
    p.keys = (unsigned long long*)(base + L.keys); p.sorted = (unsigned long long*)(base + L.sorted);
    cudaStream_t s = (cudaStream_t)stream;
    p.hist = nullptr; p.cursor = nullptr; p.grouped = nullptr;
    if (L.bucketed) {
        p.hist = (int*)(base + L.hist); p.cursor = (int*)(base + L.cursor); p.grouped = (unsigned long long*)(base + L.grouped);
        if (cudaMemsetAsync(base + L.hist, 0, L.grouped - L.hist, s) != cudaSuccess) return set_error(-EIO, "nms: memset failed");
    }
    launch_k(nms_count, dim3(dim3(L.nchunks, N)), dim3(NMS_T), 0, s, p);
    MGDT_LAUNCH_CHECK("nms_count");
    launch_k(nms_compact, dim3(dim3(L.nchunks, N)), dim3(NMS_T), 0, s, p);
    MGDT_LAUNCH_CHECK("nms_compact");
    if (L.bucketed) {
        launch_k(nms_bscan, dim3(N), dim3(1024), 0, s, p);
        MGDT_LAUNCH_CHECK("nms_bscan");
        launch_k(nms_scatter, dim3(dim3(cdiv(L.cap, NMS_T), N)), dim3(NMS_T), 0, s, p);
        MGDT_LAUNCH_CHECK("nms_scatter");
        launch_k(nms_brank, dim3(dim3(cdiv(L.cap, NMS_T), N)), dim3(NMS_T), 0, s, p);
        MGDT_LAUNCH_CHECK("nms_brank");
    } else {
        launch_k(nms_rank, dim3(dim3(cdiv(L.cap, NMS_T), N)), dim3(NMS_T), 0, s, p);
        MGDT_LAUNCH_CHECK("nms_rank");
    }
    const size_t smem = (sizeof(Box) + sizeof(float)) * (size_t)(max_det + NMS_CHUNK) +
                        sizeof(unsigned) * (size_t)(NMS_CHUNK * NMS_WORDS + NMS_WORDS);
    if (smem + 4096 > 48 * 1024) {  // the 48 KB default covers static + dynamic shared memory together
        cudaError_t e = cudaFuncSetAttribute(nms_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(-EIO, "nms: smem attr: %s", cudaGetErrorString(e));
    }
    launch_k(nms_scan, dim3(N), dim3(NMS_CHUNK), smem, s, p, out, counts);
    MGDT_LAUNCH_CHECK("nms_scan");
    return 0;
}
