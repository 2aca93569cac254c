// TMA-fed tcgen05 / TMEM kernel for the transform-free 3x3 stride-1 convolutions with Cin <= 64 (mode 1 of conv_umma2.cu).
// Included by conv_umma2.cu inside namespace mgdt, after conv_tma1x1.cuh (shares its helpers, the epilogue arithmetic and the
// fused output statistics of conv_umma2_kernel: STATS = 1 for the Conv_GN layers of the TOOD head).
//
// A operand: the same no-swizzle K-major channel planes [Cin/8][positions][16 B] as conv_umma2_kernel, positions = linear
// indices of the zero-padded image (row pitch Wq = W + 2), so a tap (dy, dx) is the plane read through a descriptor whose
// start address is shifted by (dy * Wq + dx) * 16 B.  But a tile is now R WHOLE padded image rows, and the stage of a
// tile is filled by ONE cp.async.bulk.tensor.4d (SASS UTMALDG) per 8-channel plane: box {8 channels, Wq columns, R + 2
// rows, 1 image} at (plane * 8, -1, y0 - 1, n) of the NHWC activation -- the left / right halo columns, the rows above
// and below the image and the rows past its end are zero-filled by the TMA unit (out-of-bounds fill), which is exactly
// the convolution's padding.  tools/ubench/tma_box16.cu: 1.0-1.8 TB/s for these 16-byte-inner boxes, against 11 warps
// of LDGSTS (one 512-byte instruction per ~200 cycles per warp) in conv_umma2_kernel.
//
// With the loader down to one elected lane the CTA is 6 or 10 warps (producer, MMA, 4 / 8 epilogue warps) with 45-110 KB
// of shared memory, so two CTAs share an SM: their ramps, MMA phases and epilogues interleave, and the epilogue (the
// bound of these layers once the MMAs issue from the uniform datapath) has 8-16 warps per SM.
//
//   warp 0    producer  resident weights by cp.async.bulk, then per tile expect_tx + one 4-D box per plane
//   warp 1    MMA       K = 16 steps x MB row blocks, A descriptors from the kernel parameters (p.adesc), uniform datapath
//   warps 2+  epilogue  tcgen05.ld -> bias / activation / residual / bf16 -> swizzled staging tile -> 16-byte stores
//
// Output row m of a tile (n, ty):  r = m / Wq, x = m % Wq, y = ty * R + r;  valid iff r < R, x < Wo, y < Ho.
//
// Stride 2 (same kernel, t.s2): the padded input is split into its four row / column parity sub-images while loading
// (space-to-depth, as conv_umma2_kernel's mode 2), after which every tap is again a pure shift of one sub-plane.  TMA does
// the split: the NHWC tensor is described as (channels of a COLUMN PAIR, W / 2, H, N) -- the column parity becomes a
// channel offset, so no element stride is needed along x (the widest layer has 2 * 161 > 256 columns) -- with element
// stride 2 along y.  Sub-image (ph, pw) of the tile's rows is the box {8, Wq = Wo + 1, R + 1 rows, 1} at
// (plane * 8 + (pw ? 0 : x_cs), pw ? 0 : -1, 2 * y0 + ph - 1, n): input column 2c + pw - 1 is the second pixel of pair
// c - 1 (pw = 0) or the first pixel of pair c (pw = 1).

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

static constexpr unsigned t3_tail(int epiw) { return 1024u + 256u * 4u + (unsigned)epiw * 2048u + 512u; }

template <int EPIW, int STATS>
__global__ void __launch_bounds__(64 + 32 * EPIW, 2) conv3x3_tma_kernel(const __grid_constant__ P2 p) {
    constexpr int T3_THREADS = 64 + 32 * EPIW;
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const T3& t = p.t3;
    const Plan2& pl = p.pl;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (uniform role branches)

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

    for (int i = tid; i < pl.Nc; i += T3_THREADS) sBias[i] = (p.bias && i < p.Cout) ? p.bias[i] * t1_bias_scale(p.act) : 0.f;
    // The positions of a plane behind the TMA box (never written by the loads) are read by the taps of the tile's last
    // rows -- for junk rows, and by the zero-weight dummy chunk of Cin = 8 layers for valid ones: keep them zero
    // (NaN bit patterns times zero weights would poison valid accumulators).
    {
        const uint32_t slack16 = (uint32_t)(t.Ppar - t.PB);
        const uint32_t per_stage = (uint32_t)(pl.planes * t.npar) * slack16;
        for (uint32_t i = tid; slack16 && i < (uint32_t)t.S * per_stage; i += T3_THREADS) {
            const uint32_t s = i / per_stage, r = i - s * per_stage, sub = r / slack16, q = r - sub * slack16;   // sub = plane * npar + parity
            *reinterpret_cast<uint4*>(base + t.w_bytes + (size_t)s * t.stage_bytes + ((size_t)sub * t.Ppar + t.PB + q) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)),
                     "r"((uint32_t)t.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < t.S; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(ACCFULL(a), 1); mbar_init(ACCEMPTY(a), EPIW); }
        mbar_init(WREADY, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zeroed slack is read by the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) { trace_mark(p, 0); trace_mark(p, 1); }

    const uint32_t tiles = (uint32_t)t.tiles;

    if (warp == 0) {
        // =============================================================== producer (one elected lane)
        if (elect_one()) {   // resident weights are constants of the layer: fetch them before the dependency wait
            mbar_expect_tx(WREADY, t.w_copy_bytes);
            for (uint32_t o = 0; o < t.w_copy_bytes; o += 16384u)
                bulk_load(sW32 + o, reinterpret_cast<const unsigned char*>(p.w) + o, min(16384u, t.w_copy_bytes - o), WREADY);
        }
        __syncwarp();
        pdl_wait();
        int s = 0;
        uint32_t ph = 0, tl = 0;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tl) {
            const uint32_t n_img = fdiv(tile, p.d_tpi), ty = tile - n_img * (uint32_t)t.tiles_per_img;
            mbar_wait(EMPTY(s), ph ^ 1);
            if (elect_one()) {
                const uint32_t sA = sStage32 + (uint32_t)s * t.stage_bytes;
                mbar_expect_tx(FULL(s), t.tx_bytes);
                if (!t.s2) {
                    for (int pll = 0; pll < pl.planes; ++pll)
                        tma_load_4d(sA + (uint32_t)pll * (uint32_t)t.pstride * 16u, &p.xmap, pll * 8, -1, (int)(ty * (uint32_t)t.R) - 1, (int)n_img, FULL(s));
                } else {
                    const int yb = 2 * (int)(ty * (uint32_t)t.R) - 1;
                    for (int pll = 0; pll < pl.planes; ++pll)
                        for (int par = 0; par < 4; ++par) {
                            const int ph = par >> 1, pw = par & 1;
                            tma_load_4d(sA + ((uint32_t)pll * (uint32_t)t.pstride + (uint32_t)par * (uint32_t)t.Ppar) * 16u, &p.xmap,
                                        pll * 8 + (pw ? 0 : t.xoff2), pw ? 0 : -1, yb + ph, (int)n_img, FULL(s));
                        }
                }
                if (p.trace && tl < 6) trace_mark(p, 8 + 8 * (int)tl);
            }
            __syncwarp();
            if (++s == t.S) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // =============================================================== MMA issuer (whole warp uniform, elected lane issues)
        const uint32_t idesc = (1u << 4) | (1u << 7) | ((p.w_f16 ? 0u : 1u) << 10) | ((uint32_t)(pl.Nc >> 3) << 17) | ((128u >> 4) << 24);
        const bool lead = elect_one();
        mbar_wait(WREADY, 0);
        int s = 0;
        uint32_t ph = 0, ti = 0;
        const uint64_t bdesc_t = mk_desc(0u, (uint32_t)pl.Nc * 16u, 128u) + (uint64_t)(sW32 >> 4);
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
            const int a = t.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = t.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            mbar_wait(ACCEMPTY(a), aphase ^ 1);
            mbar_wait(FULL(s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t abase = (uint64_t)((sStage32 + (uint32_t)s * t.stage_bytes) >> 4);
            const uint32_t d0 = tmem_base + (uint32_t)(a * t.MB * pl.Nc);
            // one copy of the K loop per tile height, the MB instructions of a step under ONE branch on the elected lane
            auto kloop = [&](auto mbc) {
                constexpr int MBK = decltype(mbc)::value;
                uint64_t bdesc = bdesc_t;
#pragma unroll 2
                for (int i = 0; i < pl.nmma_s; ++i, bdesc += (uint64_t)(2 * pl.Nc)) {
                    const uint64_t ad0 = p.adesc[i] + abase;
                    const uint32_t acc = i ? 1u : 0u;
                    if (lead) {
#pragma unroll
                        for (int mb = 0; mb < MBK; ++mb)   // 2048 B of A and Nc accumulator columns per row block
                            asm volatile(
                                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                                ::"r"(d0 + (uint32_t)(mb * pl.Nc)), "l"(ad0 + (uint64_t)(mb * 128)), "l"(bdesc), "r"(idesc), "r"(acc));
                    }
                }
            };
            if (t.MB == 4) kloop(std::integral_constant<int, 4>());
            else if (t.MB == 2) kloop(std::integral_constant<int, 2>());
            else kloop(std::integral_constant<int, 1>());
            if (lead) {
                umma_commit(EMPTY(s));
                umma_commit(ACCFULL(a));
                if (p.trace && ti < 6) trace_mark(p, 9 + 8 * (int)ti);
            }
            __syncwarp();
            if (++s == t.S) { s = 0; ph ^= 1; }
        }
    } else {
        // =============================================================== epilogue (TMEM quadrant = warp % 4)
        pdl_wait();
        const int ew = warp - 2, quad = warp & 3, sub = ew >> 2;
        constexpr int NSUBW = EPIW / 4;
        const int ncch = (pl.Nc + 31) / 32;
        const uint32_t stg32 = sOut32 + (uint32_t)ew * 2048u;
        const int srow = lane >> 2, schunk = lane & 3;
        const uint32_t sw_wr = (uint32_t)((lane >> 1) & 3);
        const uint32_t st_wr = stg32 + (uint32_t)lane * 64u;
        const bool trw = p.trace && ew == 0 && lane == 0;
        uint32_t ti = 0;
        for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
            const int a = t.NACC == 2 ? (int)(ti & 1) : 0;
            const uint32_t aphase = t.NACC == 2 ? ((ti >> 1) & 1) : (ti & 1);
            const uint32_t n_img = fdiv(tile, p.d_tpi), ty = tile - n_img * (uint32_t)t.tiles_per_img;
            const uint32_t y0 = ty * (uint32_t)t.R;
            mbar_wait(ACCFULL(a), aphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trw && ti < 6) trace_mark(p, 10 + 8 * (int)ti);
            int u = sub;                                           // units are numbered mb * ncch + cc
            for (int mb = 0; mb < t.MB; ++mb) {
                if (u >= (mb + 1) * ncch) continue;
                // output pixel of this lane's row
                const uint32_t m = (uint32_t)(mb * 128 + quad * 32 + lane);
                const uint32_t r = fdiv(m, p.d_Wq), x = m - r * (uint32_t)t.Wq;
                const int opix = (r < (uint32_t)t.R && x < (uint32_t)t.Wo && y0 + r < (uint32_t)t.Ho)
                                     ? (int)((n_img * (uint32_t)t.Ho + y0 + r) * (uint32_t)t.Wo + x) : -1;
                const bool any_row = __any_sync(0xffffffffu, opix >= 0);
                uint32_t skey = 0xffffffffu;   // fused statistics: (image << 4) | adaptive-pool window mask of this lane's row
                if (STATS && opix >= 0) {
                    uint32_t mask = 0;
                    if (p.st_Q == 5) {
                        const int h = (int)(y0 + r), w = (int)x;
                        const uint32_t top = h < p.st_h0e, bot = h >= p.st_h1b, lef = w < p.st_w0e, rig = w >= p.st_w1b;
                        mask = (top & lef) | ((top & rig) << 1) | ((bot & lef) << 2) | ((bot & rig) << 3);
                    }
                    skey = (n_img << 4) | mask;
                }
                __nv_bfloat16* yrow[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int orow = __shfl_sync(0xffffffffu, opix, g * 8 + srow);
                    yrow[g] = orow >= 0 ? p.y + (size_t)orow * p.y_cs : nullptr;
                }
                for (; u < (mb + 1) * ncch; u += NSUBW) {
                    const int cl = (u - mb * ncch) * 32;
                    const int co0 = cl;
                    if (!any_row || co0 >= p.Cout) continue;
                    const int nv = min(32, pl.Nc - cl);
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * t.MB * pl.Nc + mb * pl.Nc + cl);
                    uint32_t rr[32], pk[16];
                    const bool tr = trw && ti == 1;
                    long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
                    if (tr) tc0 = clock64();
                    if (nv == 32) {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
                              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]),
                              "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]),
                              "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (tr) tc1 = clock64();
                        epi_fast_rt<32>(p, rr, sBias, cl, co0, opix, pk);
                    } else {
                        asm volatile(
                            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
                              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
                            : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (tr) tc1 = clock64();
                        epi_fast_rt<16>(p, rr, sBias, cl, co0, opix, pk);
#pragma unroll
                        for (int j = 8; j < 16; ++j) pk[j] = 0u;
                    }
                    // row `lane` -> staging: 64 bytes per row, 16-byte chunk c at slot c ^ ((row >> 1) & 3)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_wr + (((uint32_t)c ^ sw_wr) << 4)), "r"(pk[4 * c]),
                                     "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                    __syncwarp();
                    if (tr) tc2 = clock64();
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
                    if (STATS)   // per-(image, channel) sums of the staged (bf16-rounded) unit, fp64 atomics (see epi_stats)
                        epi_stats(p.st_acc + (size_t)(tile % (uint32_t)p.st_R) * p.st_rs, p.st_Q, p.st_sq, p.st_tot, p.Cout, stg32, skey, lane, nv,
                                  co0, p.Cout);
                    __syncwarp();
                    if (tr) {
                        tc3 = clock64();
                        unsigned long long* tp = p.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 64 + 56;
                        tp[0] += 1; tp[1] += (unsigned long long)(tc1 - tc0); tp[2] += (unsigned long long)(tc2 - tc1);
                        tp[3] += (unsigned long long)(tc3 - tc2);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(ACCEMPTY(a));
            if (trw && ti < 6) trace_mark(p, 11 + 8 * (int)ti);
        }
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
static int g_use_tma3 = 1;      // option "conv_tma3x3": TMA-fed kernel for transform-free 3x3 layers
static int g_use_tma3_s2 = 1;   // option "conv_tma3x3_s2": ... also for stride 2

static bool plan_t3(const P2& p, T3& best) {
    const Plan2& pl = p.pl;
    const bool s2 = pl.mode == 2;
    if ((pl.mode != 1 && !s2) || pl.nks != 1 || pl.nsplit != 1 || pl.PS != pl.planes) return false;
    if (pl.planes * (s2 ? 4 : 1) > 32) return false;
    if (s2 && ((p.W & 1) || (p.H & 1))) return false;                    // column pairs; even H keeps Ho = H / 2
    const int Wq = s2 ? p.Wo + 1 : p.W + 2, Ho = p.Ho, Wo = p.Wo;
    if (Wq > 256 || Ho < 1) return false;
    const unsigned w_copy = (unsigned)pl.nmma_s * 2u * (unsigned)pl.Nc * 16u;
    int best_score = -1;
    for (int mb : {4, 2, 1}) {
        int R = std::min(128 * mb / Wq, Ho);
        if (R < 1 || 2 * (R + 2) > 256) continue;
        if (mb > 1 && (long long)R * Wq * 10 < 128LL * mb * 7) continue;   // less than 70 % of the tile's rows are pixels
        for (int epiw : {8, 4}) {
            T3 c{};
            c.MB = mb; c.R = R; c.Wq = Wq; c.epiw = epiw; c.s2 = s2 ? 1 : 0; c.npar = s2 ? 4 : 1; c.Ho = Ho; c.Wo = Wo; c.xoff2 = p.x_cs;
            c.PB = (s2 ? R + 1 : R + 2) * Wq;
            const int reach = 128 * mb + (s2 ? Wq + 2 : 2 * Wq + 3);      // furthest position a tap of the tile's last row reads
            c.Ppar = (std::max(c.PB, reach) + 7) / 8 * 8;
            c.pstride = c.npar * c.Ppar;
            c.tiles_per_img = (Ho + R - 1) / R;
            c.tiles = (long long)c.tiles_per_img * p.N;
            c.tx_bytes = (unsigned)(pl.planes * c.npar) * (unsigned)c.PB * 16u;
            c.w_copy_bytes = w_copy;
            c.w_bytes = (w_copy + 1023u) / 1024u * 1024u;
            c.stage_bytes = ((unsigned)pl.planes * (unsigned)c.pstride * 16u + 1023u) / 1024u * 1024u;
            c.NACC = 2;
            int cols = 32;
            while (cols < 2 * mb * pl.Nc) cols <<= 1;
            if (cols > 512) continue;
            c.tmem_cols = cols;
            for (int S : {3, 2}) {
                c.S = S;
                c.smem_total = c.w_bytes + (unsigned)S * c.stage_bytes + t3_tail(epiw) + 1024u;
                if (c.smem_total > (unsigned)U2_MAX_SMEM) continue;
                int ctas = std::min(std::min(512 / cols, (int)((228u * 1024u - 1024u) / (c.smem_total + 1024u))), 2);
                if (ctas < 1) continue;
                c.ctas_per_sm = ctas;
                // prefer two CTAs per SM, then enough tiles for every CTA slot, then the larger tile / more stages / warps
                const bool enough = c.tiles >= (long long)(148 * ctas) * 8 / 10;
                const int score = (ctas >= 2 ? 1000 : 0) + (enough ? 500 : 0) + mb * 20 + (epiw == 8 ? 8 : 0) + S;
                if (score > best_score) { best_score = score; best = c; }
            }
        }
    }
    return best_score >= 0;
}

static void fill_adesc3(P2& p) {
    const Plan2& pl = p.pl;
    const T3& t = p.t3;
    auto off = [&](int c) -> uint32_t {
        const int tp = c / pl.PS, pll = c - tp * pl.PS;
        if (t.s2)   // parity sub-plane of the tap, shifted by (dy >> 1, dx >> 1)
            return ((uint32_t)pll * (uint32_t)t.pstride + (uint32_t)pl.tap_par[tp] * (uint32_t)t.Ppar +
                    (uint32_t)((pl.tap_dy[tp] >> 1) * t.Wq + (pl.tap_dx[tp] >> 1))) * 16u;
        return ((uint32_t)pll * (uint32_t)t.pstride + (uint32_t)(pl.tap_dy[tp] * t.Wq + pl.tap_dx[tp])) * 16u;
    };
    for (int i = 0; i < pl.nmma_s && i < U2_MAX_MMA; ++i) {
        const int c0 = 2 * i, c1 = 2 * i + 1;
        const uint32_t o0 = off(c0);
        const uint32_t lbo = (c1 < pl.taps * pl.PS) ? (off(c1) - o0) : 16u;  // dummy chunk: its weights are zero
        p.adesc[i] = (uint64_t)((o0 >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
    }
}

template <int EPIW, int STATS>
static int launch_t3k(const P2& p, dim3 grid, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tma_kernel<EPIW, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, U2_MAX_SMEM);
    if (e != cudaSuccess) return set_error(-EIO, "conv3x3_tma: smem attr: %s", cudaGetErrorString(e));
    launch_k(conv3x3_tma_kernel<EPIW, STATS>, grid, dim3(64 + 32 * EPIW), p.t3.smem_total, s, p);
    MGDT_LAUNCH_CHECK("conv3x3_tma");
    return 0;
}

static bool t3_eligible(const P2& p) {
    return g_use_tma3 && (p.pl.mode == 1 || (p.pl.mode == 2 && g_use_tma3_s2)) && !p.stem_src && !p.dcn_off && !p.pre_add && !p.in_scale && !p.pix_scale &&
           !p.in_relu && !p.row_scale && !p.act_cols && !p.w_img_elems && p.Cout <= 256;
}

// Try the TMA kernel for this (already filled) mode-1 layer: 1 = launched, 0 = not eligible, < 0 = error.
static int try_launch_t3(P2& p, cudaStream_t s) {
    if (!t3_eligible(p)) return 0;
    if (!plan_t3(p, p.t3)) return 0;
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return 0;
    const T3& t = p.t3;
    cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)p.x_cs * 2, (cuuint64_t)p.W * p.x_cs * 2, (cuuint64_t)p.H * p.W * p.x_cs * 2};
    cuuint32_t box[4] = {8, (cuuint32_t)t.Wq, (cuuint32_t)(t.R + 2), 1}, estr[4] = {1, 1, 1, 1};
    if (t.s2) {   // column pairs as channels, element stride 2 along y: R + 1 rows of one row parity per box
        dims[0] = (cuuint64_t)p.x_cs + (cuuint64_t)p.Cin; dims[1] = (cuuint64_t)(p.W / 2);
        strides[0] = (cuuint64_t)p.x_cs * 4;
        box[2] = (cuuint32_t)(2 * (t.R + 1)); estr[2] = 2;
    }
    if (encode(&p.xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)p.x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    p.d_tpi = make_fastdiv((uint32_t)t.tiles_per_img);
    p.d_Wq = make_fastdiv((uint32_t)t.Wq);
    fill_adesc3(p);
    const long long cx = std::min<long long>(t.tiles, 148LL * t.ctas_per_sm);
    const dim3 grid((unsigned)cx, 1);
    const int rc = t.epiw == 8 ? (p.st_acc ? launch_t3k<8, 1>(p, grid, s) : launch_t3k<8, 0>(p, grid, s))
                               : (p.st_acc ? launch_t3k<4, 1>(p, grid, s) : launch_t3k<4, 0>(p, grid, s));
    return rc < 0 ? rc : 1;
}
