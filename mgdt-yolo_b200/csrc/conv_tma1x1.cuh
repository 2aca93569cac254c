// TMA-fed tcgen05 / TMEM kernel for the transform-free 1x1 convolutions (mode 0 of conv_umma2.cu).
// Included by conv_umma2.cu inside namespace mgdt, after the shared epilogue helpers (epi_math / epi_stats).
//
// Why a second kernel: in conv_umma2_kernel the A operand (pixels x Cin) is gathered by 3-11 warps of 16-byte cp.async
// (LDGSTS) into channel planes; for a 1x1 layer that gather is a plain 2-D tile of the NHWC activation, which ONE TMA
// instruction per 64-channel block loads (cp.async.bulk.tensor.2d, SASS UTMALDG) straight into the canonical
// 128B / 64B / 32B-swizzled K-major UMMA layout.  That frees the LSU pipeline and 10 warps, so the CTA shrinks to six
// warps (one TMA lane, one MMA lane, four epilogue warps = the four TMEM lane quadrants) with 45-110 KB of shared
// memory: two to four CTAs are resident per SM, each with its own pipeline, and the epilogue (the bound of these
// layers) has 8-16 independent warps per SM instead of two per sub-partition fighting the producers for issue slots.
//
//   warp 0   producer  elect.sync lane: mbarrier.arrive.expect_tx(full[s]) + one cp.async.bulk.tensor.2d per K block
//                      (+ cp.async.bulk copies of the matching weight slice when the weights change per image)
//   warp 1   MMA       owns the TMEM allocation; elect.sync lane issues tcgen05.mma.kind::f16 (A, B from shared
//                      memory through swizzled / plain K-major descriptors), tcgen05.commit -> empty[s] / accfull[a]
//   warps 2-5 epilogue tcgen05.ld 32x32b -> bias / activation / residual / bf16 -> swizzled staging tile -> TMA store
//                      (UTMASTG; two staging tiles per warp so the store of unit u overlaps the math of unit u + 1)
//
// Shared memory:  [resident weights] [S ring stages, 1024-byte aligned: A K-blocks (+ weight slice)] [barriers | TMEM
// slot | bias] [epilogue staging 4 warps x 2 x 2 KB, 512-byte aligned].
// Weights keep the packed image of mgdt_conv_umma_pack ([chunk of 8 input channels][Nc][8], no-swizzle K-major);
// a CTA copies only its Nsub columns of every chunk, so column sub-splits need no repacking.

constexpr int T1_MAX_STAGES = 4;
// barriers + TMEM slot (1 KB), bias (1 KB), two 2 KB staging tiles per epilogue warp (512-byte aligned)
static constexpr unsigned t1_tail(int epiw) { return 1024u + 256u * 4u + (unsigned)epiw * 2u * 2048u + 512u + (unsigned)epiw * 2u * 6u * 32u * 4u; }

__device__ __forceinline__ uint64_t mk_desc_sw(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout) {
    // K-major swizzled shared-memory matrix descriptor (version 1): LBO field = 1 (unused for swizzled K-major),
    // SBO = byte distance of consecutive 8-row groups, layout 2 / 4 / 6 = SWIZZLE_128B / 64B / 32B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// EPIW = 4: six warps, three CTAs per SM (narrow layers: one 32-column unit per TMEM quadrant and tile);
// EPIW = 8: ten warps, two CTAs per SM, the two warps of a quadrant take the tile's column units alternately.
// Fused output statistics, TMA kernel form.  Same per-unit reduction as epi_stats (lane l owns the column pair
// 2 * (l & 15) over the 16 rows 2i + (l >> 4) of the staged unit, one shuffle combines the half-warps), but the results
// are added to PER-WARP fp32 accumulators in shared memory (acc[unit slot][plane][32 columns], one owner lane per
// address: no atomics) that are flushed to the global fp64 accumulators only when the image changes or the CTA ends
// (t1_stats_flush).  With contiguous tile ranges per CTA that is once or twice per CTA instead of once per unit; the
// flush order does not matter (fp64 atomics), the per-warp sums are formed in a fixed order: reproducible results.
__device__ __forceinline__ void t1_stats_flush(const P2& p, float* wacc, int slots, int n_cur, int lane, int co_base, int slot_stride_cols) {
    if (n_cur < 0) return;
    const int K = p.st_Q + p.st_sq;
    double* base = p.st_acc + (size_t)(blockIdx.x % (unsigned)p.st_R) * p.st_rs + (size_t)n_cur * K * p.Cout;
    for (int sl = 0; sl < slots; ++sl) {
        const int c = co_base + sl * slot_stride_cols + lane;
        for (int k = 0; k < K; ++k) {
            float* a = wacc + (sl * 6 + k) * 32 + lane;
            const float v = *a;
            *a = 0.f;
            if (c < p.Cout && v != 0.f) atomicAdd(base + (size_t)k * p.Cout + c, (double)v);
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void t1_stats_unit(const P2& p, float* wacc, int slots, int& n_cur, int slot, uint32_t stg32, uint32_t skey,
                                              int lane, int co_base, int slot_stride_cols) {
    const uint32_t full = 0xffffffffu, INVALID = 0xffffffffu;
    const uint32_t cp = (uint32_t)lane & 15u, half = (uint32_t)lane >> 4;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t a = stg32 + (uint32_t)(2 * i) * 64u + half * 64u + (((cp >> 2) ^ ((uint32_t)i & 3u)) << 4) + (cp & 3u) * 4u;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[i]) : "r"(a));
    }
    const int Q = p.st_Q;
    float* my = wacc + slot * 6 * 32 + 2 * (int)cp;
    uint32_t rem = __ballot_sync(full, skey != INVALID);
    while (rem) {
        const uint32_t k = __shfl_sync(full, skey, __ffs((int)rem) - 1);
        const uint32_t m = __ballot_sync(full, skey == k);
        rem &= ~m;
        if ((int)(k >> 4) != n_cur) {   // uniform: a new image starts -- publish the finished one first
            t1_stats_flush(p, wacc, slots, n_cur, lane, co_base, slot_stride_cols);
            n_cur = (int)(k >> 4);
        }
        const uint32_t mh = m >> half;
        float2 sa[4], qa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) sa[u] = qa[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t tt = (m == full || (mh & (1u << (2 * i)))) ? w[i] : 0u;
            const float2 v2 = make_float2(__uint_as_float(tt << 16), __uint_as_float(tt & 0xffff0000u));
            sa[i & 3] = __fadd2_rn(v2, sa[i & 3]);
            qa[i & 3] = __ffma2_rn(v2, v2, qa[i & 3]);
        }
        float2 s2 = __fadd2_rn(__fadd2_rn(sa[0], sa[1]), __fadd2_rn(sa[2], sa[3]));
        float2 q2 = __fadd2_rn(__fadd2_rn(qa[0], qa[1]), __fadd2_rn(qa[2], qa[3]));
        s2.x += __shfl_xor_sync(full, s2.x, 16); s2.y += __shfl_xor_sync(full, s2.y, 16);
        q2.x += __shfl_xor_sync(full, q2.x, 16); q2.y += __shfl_xor_sync(full, q2.y, 16);
        if (half == 0) {
            if (p.st_tot) { float2* a = reinterpret_cast<float2*>(my); *a = __fadd2_rn(*a, s2); }
            if (Q == 5) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (k & (1u << j)) { float2* a = reinterpret_cast<float2*>(my + (1 + j) * 32); *a = __fadd2_rn(*a, s2); }
            }
            if (p.st_sq) { float2* a = reinterpret_cast<float2*>(my + Q * 32); *a = __fadd2_rn(*a, q2); }
        }
        __syncwarp();
    }
}

template <int STATS, int EPIW>
__global__ void __launch_bounds__(64 + 32 * EPIW, EPIW == 4 ? 3 : 2) conv1x1_tma_kernel(const __grid_constant__ P2 p) {
    constexpr int T1_THREADS = 64 + 32 * EPIW, T1_EPI_WARPS = EPIW;
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const T1& t = p.t1;
    const Plan2& pl = p.pl;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform: uniform role branches, uniform-datapath descriptor arithmetic
    const int ns = (int)blockIdx.y / t.nsub, hs = (int)blockIdx.y - ns * t.nsub;   // packed column block, sub-split
    const int col0 = hs * t.Nsub;                                                  // first column within the block
    const int Nsub = t.Nsub;

    // the dynamic segment is only 16-byte aligned by contract: align the ring by hand (swizzle atoms repeat every 1 KB)
    const uint32_t base32 = (s_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* base = smem_raw + (base32 - s_u32(smem_raw));
    const uint32_t sW32 = base32;                                   // resident weights (w_bytes, a multiple of 1024)
    const uint32_t sStage32 = base32 + t.w_bytes;
    unsigned char* tail = base + t.w_bytes + (size_t)t.S * t.stage_bytes;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tail);
    const uint32_t bar0 = s_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (T1_MAX_STAGES + s); };
    auto ACCFULL = [&](int a) { return bar0 + 8u * (2 * T1_MAX_STAGES + a); };
    auto ACCEMPTY = [&](int a) { return bar0 + 8u * (2 * T1_MAX_STAGES + 2 + a); };
    const uint32_t WREADY = bar0 + 8u * (2 * T1_MAX_STAGES + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T1_MAX_STAGES + 5);
    float* sBias = reinterpret_cast<float*>(tail + 1024);
    const uint32_t sOut32 = (s_u32(tail + 1024 + 256 * 4) + 511u) & ~511u;
    // per-warp statistics accumulators (STATS): [EPIW warps][2 unit slots][6 planes][32 columns] fp32, after the staging tiles
    float* sStat = reinterpret_cast<float*>(tail + 1024 + 256 * 4 + 512 + (size_t)EPIW * 4096);
    if (STATS) {
        for (int i = tid; i < EPIW * 2 * 6 * 32; i += T1_THREADS) sStat[i] = 0.f;
    }

    for (int i = tid; i < Nsub; i += T1_THREADS) {
        const int co = ns * pl.Nc + col0 + i;
        sBias[i] = (p.bias && co < p.Cout) ? p.bias[co] * t1_bias_scale(p.act) : 0.f;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                     "r"((uint32_t)t.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < t.S; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACCFULL(a), 1); mbar_init(ACCEMPTY(a), T1_EPI_WARPS); }
        mbar_init(WREADY, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) { trace_mark(p, 0); trace_mark(p, 1); }

    // contiguous tile range per CTA: consecutive tiles mostly belong to one image, so the fused statistics are
    // accumulated per warp in shared memory and flushed to the global fp64 accumulators only when the image changes
    const uint32_t tiles_all = (uint32_t)t.tiles;
    const uint32_t tchunk = (tiles_all + gridDim.x - 1) / gridDim.x;
    const uint32_t tile_lo = min(tiles_all, blockIdx.x * tchunk), tiles = min(tiles_all, tile_lo + tchunk);
    const uint32_t HW = (uint32_t)(p.H * p.W);
    const int nchunk_all = p.Cin / 8;                       // 16-byte K chunks of the packed weight image
    const uint32_t wchunk_bytes = (uint32_t)Nsub * 16u;     // one chunk of this CTA's columns in shared memory
    // this CTA's columns of chunk c of image `img` in the packed image (chunks of a column block are contiguous:
    // [ns][chunk][Nc][8] for even plane counts per slice, which the host checks)
    auto wsrc = [&](uint32_t img, int c) -> const __nv_bfloat16* {
        return p.w + (size_t)img * p.w_img_elems + (((size_t)ns * nchunk_all + c) * pl.Nc + col0) * 8;
    };
    // first pixel (row of the 2-D activation view) of a tile, and the image it belongs to when tiles are cut per image
    auto tile_pix0 = [&](uint32_t tile, uint32_t& n_img) -> uint32_t {
        if (t.per_img) {
            n_img = fdiv(tile, p.d_tpi);
            return n_img * HW + (tile - n_img * (uint32_t)t.tiles_per_img) * 128u;
        }
        n_img = 0;
        return tile * 128u;
    };

    if (warp == 0) {
        // =============================================================== producer (one elected lane)
        if (!t.w_ring) {   // resident weights are constants of the layer: fetch them before the dependency wait
            if (elect_one()) {
                mbar_expect_tx(WREADY, (uint32_t)nchunk_all * wchunk_bytes);
                for (int c = 0; c < nchunk_all; ++c) bulk_load(sW32 + (uint32_t)c * wchunk_bytes, wsrc(0, c), wchunk_bytes, WREADY);
            }
            __syncwarp();
        }
        pdl_wait();
        int s = 0;
        uint32_t ph = 0;
        for (uint32_t tile = tile_lo; tile < tiles; ++tile) {
            uint32_t n_img;
            const uint32_t pix0 = tile_pix0(tile, n_img);
            for (int st = 0; st < t.nst; ++st) {
                mbar_wait(EMPTY(s), ph ^ 1);
                if (elect_one()) {
                    const int kb0 = st * t.kb_stage, nk = min(t.kb_stage, t.nkb - kb0);
                    const uint32_t sA = sStage32 + (uint32_t)s * t.stage_bytes;
                    const int cpk = t.KB / 8;   // weight chunks per K block
                    mbar_expect_tx(FULL(s), (uint32_t)nk * t.a_kb_bytes + (t.w_ring ? (uint32_t)(nk * cpk) * wchunk_bytes : 0u));
                    for (int j = 0; j < nk; ++j)
                        tma_load_2d(sA + (uint32_t)j * t.a_kb_bytes, &p.xmap, (kb0 + j) * t.KB, (int)pix0, FULL(s));
                    if (t.w_ring) {
                        const uint32_t sWs = sA + t.a_stage_bytes;
                        for (int c = 0; c < nk * cpk; ++c)
                            bulk_load(sWs + (uint32_t)c * wchunk_bytes, wsrc(n_img, kb0 * cpk + c), wchunk_bytes, FULL(s));
                    }
                    if (p.trace && st == t.nst - 1) { const uint32_t tl = tile - tile_lo; if (tl < 6) trace_mark(p, 8 + 8 * (int)tl); }
                }
                __syncwarp();
                if (++s == t.S) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer (one elected lane)
        // instruction descriptor: D = f32, A = B = bf16 (or f16 weights), both K-major, N = Nsub, M = 128
        const uint32_t idesc = (1u << 4) | (1u << 7) | ((p.w_f16 ? 0u : 1u) << 10) | ((uint32_t)(Nsub >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sbo = 8u * (uint32_t)t.KB * 2u;                       // eight rows of KB bf16
        const uint32_t lay = t.swz == 3 ? 2u : (t.swz == 2 ? 4u : 6u);
        if (!t.w_ring) mbar_wait(WREADY, 0);
        int s = 0;
        uint32_t ph = 0, ti = 0;
        // whole warp in uniform control flow, descriptors on the uniform datapath, only the tcgen05 instructions under the
        // elect.sync predicate (see conv_umma2_kernel's MMA role / tools/ubench/umma_issue.cu)
        const bool lead = elect_one();
        const uint64_t adesc_t = mk_desc_sw(0u, sbo, lay), bdesc_t = mk_desc(0u, wchunk_bytes, 128u);
        const int steps = t.KB / 16;
        for (uint32_t tile = tile_lo; tile < tiles; ++tile, ++ti) {
            const int a = t.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = t.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            mbar_wait(ACCEMPTY(a), aphase ^ 1);
            for (int st = 0; st < t.nst; ++st) {
                mbar_wait(FULL(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int kb0 = st * t.kb_stage, nk = min(t.kb_stage, t.nkb - kb0);
                const uint32_t sA = sStage32 + (uint32_t)s * t.stage_bytes;
                const uint32_t wb = t.w_ring ? sA + t.a_stage_bytes : sW32 + (uint32_t)(kb0 * (t.KB / 8)) * wchunk_bytes;
                const uint32_t d0 = tmem_base + (uint32_t)(a * t.KS * Nsub);
                uint64_t bdesc = bdesc_t + (uint64_t)(wb >> 4);
                for (int j = 0; j < nk; ++j) {
                    uint64_t adesc = adesc_t + (uint64_t)((sA + (uint32_t)j * t.a_kb_bytes) >> 4);
                    for (int e = 0; e < steps; ++e, adesc += 2, bdesc += (uint64_t)(2u * (wchunk_bytes >> 4))) {   // 32 B of K per step
                        const int g = (kb0 + j) * steps + e;               // K = 16 step of this tile
                        const uint32_t acc = g >= t.KS ? 1u : 0u;          // the first step of every partial accumulator overwrites
                        if (lead)
                            asm volatile(
                                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                ::"r"(d0 + (uint32_t)((g & (t.KS - 1)) * Nsub)), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc));
                    }
                }
                if (lead) {
                    umma_commit(EMPTY(s));
                    if (st == t.nst - 1) { umma_commit(ACCFULL(a)); if (p.trace && ti < 6) trace_mark(p, 9 + 8 * (int)ti); }
                }
                __syncwarp();
                if (++s == t.S) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // =============================================================== epilogue (warps 2-5 = TMEM quadrants 2, 3, 0, 1)
        pdl_wait();
        const int ew = warp - 2, quad = warp & 3, sub = ew >> 2;
        constexpr int NSUBW = EPIW / 4;
        const int ncch = (Nsub + 31) / 32;
        const uint32_t stg_base = sOut32 + (uint32_t)ew * 4096u;     // two 2 KB staging tiles
        const int srow = lane >> 2, schunk = lane & 3;
        const uint32_t sw_wr = (uint32_t)((lane >> 1) & 3);
        const bool tma_on = p.tma_store != 0;
        uint32_t ti = 0, ubuf = 0;
        float* wacc = sStat + ew * (2 * 6 * 32);
        int st_ncur = -1;                              // image whose partial sums the warp's accumulators hold
        const int st_slots = (ncch - sub + NSUBW - 1) / NSUBW;   // units this warp handles per tile (<= 2)
        const int st_co_base = ns * pl.Nc + col0 + sub * 32;
        for (uint32_t tile = tile_lo; tile < tiles; ++tile, ++ti) {
            const int a = t.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = t.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            uint32_t n_img;
            const uint32_t pix0 = tile_pix0(tile, n_img);
            // output pixel of this lane's row (row = quad * 32 + lane), -1 beyond the image / the tensor
            const uint32_t m = (uint32_t)(quad * 32 + lane);
            int opix;
            if (t.per_img) {
                const uint32_t q = pix0 - n_img * HW + m;
                opix = q < HW ? (int)(pix0 + m) : -1;
            } else {
                opix = pix0 + m < p.M_total ? (int)(pix0 + m) : -1;
            }
            const bool any_row = __any_sync(0xffffffffu, opix >= 0);
            uint32_t skey = 0xffffffffu;
            if (STATS && opix >= 0) {
                const uint32_t n = fdiv((uint32_t)opix, p.d_oHW);
                uint32_t mask = 0;
                if (p.st_Q == 5) {
                    const uint32_t rem = (uint32_t)opix - n * HW;
                    const int h = (int)fdiv(rem, p.d_oW), w = (int)rem - h * p.Wo;
                    const uint32_t top = h < p.st_h0e, bot = h >= p.st_h1b, lef = w < p.st_w0e, rig = w >= p.st_w1b;
                    mask = (top & lef) | ((top & rig) << 1) | ((bot & lef) << 2) | ((bot & rig) << 3);
                }
                skey = (n << 4) | mask;
            }
            __nv_bfloat16* yrow[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int orow = __shfl_sync(0xffffffffu, opix, g * 8 + srow);
                yrow[g] = orow >= 0 ? p.y + (size_t)orow * p.y_cs : nullptr;
            }
            mbar_wait(ACCFULL(a), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const bool trw = p.trace && ew == 0 && lane == 0;
            if (trw && ti < 6) trace_mark(p, 10 + 8 * (int)ti);
            for (int cc = sub; cc < ncch; cc += NSUBW) {
                const int cl = cc * 32;
                const int co0 = ns * pl.Nc + col0 + cl;
                if (!any_row || co0 >= p.Cout) continue;
                const int nv = min(32, Nsub - cl);
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * t.KS * Nsub + cl);
                uint32_t r[32], pk[16];
                const bool tr = trw && ti == 1;
                long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
                if (tr) tc0 = clock64();
                if (nv == 32) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    for (int h = 1; h < t.KS; ++h) {
                        tmem_ld_add16(taddr + (uint32_t)(h * Nsub), r);
                        tmem_ld_add16(taddr + (uint32_t)(h * Nsub + 16), r + 16);
                    }
                    if (tr) tc1 = clock64();
                    epi_fast_rt<32>(p, r, sBias, cl, co0, opix, pk, false, -1, (p.act_cols && co0 >= p.act_cols) ? MGDT_ACT_NONE : p.act);
                } else {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    for (int h = 1; h < t.KS; ++h) tmem_ld_add16(taddr + (uint32_t)(h * Nsub), r);
                    epi_fast_rt<16>(p, r, sBias, cl, co0, opix, pk, false, -1, (p.act_cols && co0 >= p.act_cols) ? MGDT_ACT_NONE : p.act);
#pragma unroll
                    for (int j = 8; j < 16; ++j) pk[j] = 0u;
                }
                const bool tma = tma_on && nv == 32;     // a 16-column tail keeps the LSU stores (the box would spill over)
                const uint32_t stg32 = stg_base + ubuf * 2048u;
                ubuf ^= 1u;
                if (tma_on) {   // the TMA store that last read THIS staging tile (two units ago) has finished reading it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                }
                // row `lane` -> staging: 64 bytes per row, 16-byte chunk c at slot c ^ ((row >> 1) & 3)  (SWIZZLE_64B)
                const uint32_t st_wr = stg32 + (uint32_t)lane * 64u;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_wr + (((uint32_t)c ^ sw_wr) << 4)), "r"(pk[4 * c]),
                                 "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                if (tma) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (tr) tc2 = clock64();
                if (tma) {
                    if (lane == 0) {
                        const uint32_t r0 = (uint32_t)(quad * 32);
                        if (t.per_img) {
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                         ::"l"(&p.ymap), "r"(co0), "r"(pix0 - n_img * HW + r0), "r"(n_img), "r"(stg32) : "memory");
                        } else {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                         ::"l"(&p.ymap), "r"(co0), "r"(pix0 + r0), "r"(stg32) : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {
                    if (tma_on && lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // keep the group count per unit
                    const int c8 = co0 + schunk * 8;
                    const bool chunk_on = schunk * 8 < nv && c8 < p.Cout;
                    const bool full8 = c8 + 8 <= p.Cout && p.y_vec;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int row = g * 8 + srow;
                        uint4 o;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w)
                                     : "r"(stg32 + (uint32_t)row * 64u + (((uint32_t)schunk ^ (uint32_t)((row >> 1) & 3)) << 4)));
                        if (yrow[g] != nullptr && chunk_on) {
                            __nv_bfloat16* yp = yrow[g] + c8;
                            if (full8) *reinterpret_cast<uint4*>(yp) = o;
                            else {
                                const __nv_bfloat16* oh = reinterpret_cast<const __nv_bfloat16*>(&o);
                                for (int j = 0; j < 8 && c8 + j < p.Cout; ++j) yp[j] = oh[j];
                            }
                        }
                    }
                }
                if (STATS) t1_stats_unit(p, wacc, st_slots, st_ncur, (cc - sub) / NSUBW, stg32, skey, lane, st_co_base, NSUBW * 32);
                __syncwarp();
                if (tr) {
                    tc3 = clock64();
                    unsigned long long* tp = p.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + 56;
                    tp[0] += 1; tp[1] += (unsigned long long)(tc1 - tc0); tp[2] += (unsigned long long)(tc2 - tc1);
                    tp[3] += (unsigned long long)(tc3 - tc2);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(ACCEMPTY(a));
            if (trw && ti < 6) trace_mark(p, 11 + 8 * (int)ti);
        }
        if (STATS) t1_stats_flush(p, wacc, st_slots, st_ncur, lane, st_co_base, NSUBW * 32);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all TMA stores complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) trace_mark(p, 3);
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)t.tmem_cols)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------- host side
// Plan of the TMA kernel for a mode-0 layer, or false if the layer stays on conv_umma2_kernel.
static bool plan_t1(const P2& p, T1& t) {
    const Plan2& pl = p.pl;
    if (pl.mode != 0 || p.Cin % 16 != 0 || p.Cin < 16) return false;
    if (pl.nks > 1 && (pl.PS & 1)) return false;        // packed chunks of a column block must be contiguous
    if ((pl.taps * pl.PS) & 1) return false;
    t.KB = p.Cin % 64 == 0 ? 64 : (p.Cin % 32 == 0 ? 32 : 16);
    t.swz = t.KB == 64 ? 3 : (t.KB == 32 ? 2 : 1);
    t.nkb = p.Cin / t.KB;
    t.a_kb_bytes = 128u * (unsigned)t.KB * 2u;
    // columns per CTA: at most 128, so that two accumulator buffers fit 256 TMEM columns and two CTAs share an SM
    t.Nsub = pl.Nc;
    t.nsub = 1;
    if (pl.Nc > 128 && (pl.Nc / 2) % 32 == 0) { t.Nsub = pl.Nc / 2; t.nsub = 2; }
    if (t.Nsub % 16 != 0 || t.Nsub > 256) return false;
    t.epiw = (t.Nsub + 31) / 32 >= 2 ? 8 : 4;
    if (p.st_acc && ((t.Nsub + 31) / 32 + t.epiw / 4 - 1) / (t.epiw / 4) > 2) return false;   // two statistics slots per warp
    const int want_ctas = t.epiw == 8 ? 2 : 3;
    // accumulator buffers: two when `want_ctas` CTAs still fit the SM's 512 TMEM columns, else one (the CTAs of an SM
    // interleave their MMA and epilogue phases instead)
    t.NACC = 2;
    int cols = 32;
    while (cols < t.NACC * t.Nsub) cols <<= 1;
    if (cols * want_ctas > 512) {
        t.NACC = 1;
        cols = 32;
        while (cols < t.Nsub) cols <<= 1;
    }
    // K-split partial accumulators (see try_run2): a tile here is one 128-row block, so its K = 16 steps form a single
    // dependent chain (~310 cycles each) unless they rotate over KS accumulators; wide-K layers (256 -> 64: 16 steps)
    // would otherwise spend 5,000 cycles per tile waiting on themselves
    t.KS = 1;
    const int ksteps = p.Cin / 16;
    if (g_ksplit)
        for (int ks = 4; ks >= 2; ks >>= 1) {
            int c2 = 32;
            while (c2 < t.NACC * ks * t.Nsub) c2 <<= 1;
            if (c2 * want_ctas <= 512 && ks * 2 <= ksteps) { t.KS = ks; cols = c2; break; }
        }
    t.tmem_cols = cols;
    t.per_img = p.w_img_elems ? 1 : 0;
    const long long HW = (long long)p.H * p.W;
    t.tiles_per_img = t.per_img ? (int)((HW + 127) / 128) : 1;
    t.tiles = t.per_img ? (long long)t.tiles_per_img * p.N : ((long long)p.M_total + 127) / 128;
    const unsigned w_all = (unsigned)(p.Cin / 8) * (unsigned)t.Nsub * 16u;
    // two candidate layouts: weights resident (copied once per CTA) or carried through the ring with every K slice
    // (mandatory for per-image weights); prefer the one that lets more CTAs share an SM, then resident
    int best_ctas = 0;
    T1 best = t;
    for (int ring = t.per_img ? 1 : 0; ring <= 1; ++ring) {
        T1 c = t;
        c.w_ring = ring;
        const unsigned w_per_kb = (unsigned)(c.KB / 8) * (unsigned)c.Nsub * 16u;
        const unsigned per_kb = c.a_kb_bytes + (ring ? w_per_kb : 0u);
        c.kb_stage = std::max(1, std::min(c.nkb, (int)(32768u / per_kb)));
        c.nst = (c.nkb + c.kb_stage - 1) / c.kb_stage;
        c.a_stage_bytes = (unsigned)c.kb_stage * c.a_kb_bytes;
        c.w_stage_bytes = ring ? (unsigned)c.kb_stage * w_per_kb : 0u;
        c.stage_bytes = (c.a_stage_bytes + c.w_stage_bytes + 1023u) & ~1023u;
        c.w_bytes = ring ? 0u : ((w_all + 1023u) & ~1023u);
        c.S = std::min(T1_MAX_STAGES, std::max(2, std::min(c.nst * 3, 3)));
        c.smem_total = 1024u + c.w_bytes + (unsigned)c.S * c.stage_bytes + t1_tail(c.epiw);
        if (c.smem_total > (unsigned)U2_MAX_SMEM) continue;
        int ctas = std::min(std::min(512 / c.tmem_cols, (int)((228u * 1024u - 1024u) / (c.smem_total + 1024u))), want_ctas);
        if (ctas < 1) continue;
        c.ctas_per_sm = ctas;
        if (ctas > best_ctas) { best_ctas = ctas; best = c; }
    }
    if (!best_ctas) return false;
    t = best;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    // resolved at run time: the library must load on machines without libcuda (host-side checks only)
    static EncodeTiledFn encode = nullptr;
    static bool looked_up = false;
    if (!looked_up) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = (EncodeTiledFn)fn;
        (void)cudaGetLastError();
        looked_up = true;
    }
    return encode;
}

template <int STATS, int EPIW>
static int launch_t1k(const P2& p, dim3 grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(conv1x1_tma_kernel<STATS, EPIW>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_MAX_SMEM);
    if (e != cudaSuccess) return set_error(-EIO, "conv1x1_tma: smem attr: %s", cudaGetErrorString(e));
    launch_k(conv1x1_tma_kernel<STATS, EPIW>, grid, dim3(64 + 32 * EPIW), p.t1.smem_total, s, p);
    MGDT_LAUNCH_CHECK("conv1x1_tma");
    return 0;
}

// Try the TMA kernel for this (already filled) mode-0 layer: 1 = launched, 0 = not eligible, < 0 = error.
static bool t1_eligible(const P2& p) {
    if (p.act_cols && t1_bias_scale(p.act) != 1.0f) return false;   // the bias is pre-scaled per launch, not per column range
    return g_use_tma_loads && (g_tma_stats || !p.st_acc) && !p.stem_src && !p.dcn_off && !p.pre_add && !p.in_scale && !p.pix_scale && !p.in_relu;
}

static int try_launch_t1(P2& p, cudaStream_t s) {
    if (!t1_eligible(p)) return 0;
    if (!plan_t1(p, p.t1)) return 0;
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return 0;
    const T1& t = p.t1;
    p.d_tpi = make_fastdiv((uint32_t)t.tiles_per_img);     // this kernel's own tiling (128-row tiles)
    cuuint64_t dims[2] = {(cuuint64_t)p.Cin, (cuuint64_t)p.M_total};
    cuuint64_t strides[1] = {(cuuint64_t)p.x_cs * 2};
    cuuint32_t box[2] = {(cuuint32_t)t.KB, 128}, estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = t.swz == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : (t.swz == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    if (encode(&p.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    const int ny = p.pl.nsplit * t.nsub;
    long long cx = std::min<long long>(t.tiles, std::max<long long>(1, (148LL * t.ctas_per_sm) / ny));
    const dim3 grid((unsigned)cx, (unsigned)ny);
    const int rc = t.epiw == 8 ? (p.st_acc ? launch_t1k<1, 8>(p, grid, s) : launch_t1k<0, 8>(p, grid, s))
                               : (p.st_acc ? launch_t1k<1, 4>(p, grid, s) : launch_t1k<0, 4>(p, grid, s));
    return rc < 0 ? rc : 1;
}
