/*
 * mgdt_b200.h -- C ABI of libmgdt_b200.so, the B200 (sm_100a) kernels behind the
 * MGDT-YOLO detection forward path.
 *
 * The reference (zzuiekongning/MGDT-YOLO, an Ultralytics 8.0.120 fork) is pure
 * Python and has no FFI of its own; its "operator
 * interface" for this path is the nn.Module class surface of nn/modules/ plus
 * yolo/utils/ops.py:non_max_suppression.  Each entry point below names the
 * reference op(s) it replaces (paths relative to the reference root).  The
 * Python modules in mgdt-yolo_b200/modules/ keep the reference's class names,
 * constructor/forward signatures and state_dict keys and call ONLY these
 * functions for arithmetic (ctypes).
 *
 * Conventions
 *   - caller owns all memory; the library never allocates device memory and
 *     keeps no pointer after return;
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as void*),
 *     nothing synchronises; every function is CUDA-graph capturable;
 *   - activations are NHWC ("channels_last") with an explicit channel stride
 *     `*_cs` (elements between consecutive pixels), so a channel slice of a
 *     concat buffer is addressed without a copy;
 *   - dtype: MGDT_F32 (validation mode) or MGDT_BF16 (bf16 storage, fp32 accumulate);
 *   - return 0 on success, negative errno-style code on failure; the message is
 *     available (thread-local) from mgdt_last_error().
 */
#ifndef MGDT_B200_H
#define MGDT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGDT_ABI_VERSION 5

enum { MGDT_F32 = 0, MGDT_BF16 = 1 };
enum { MGDT_ACT_NONE = 0, MGDT_ACT_SILU = 1, MGDT_ACT_RELU = 2, MGDT_ACT_SIGMOID = 3, MGDT_ACT_HSIGMOID = 4,
       MGDT_ACT_GELU = 5 };
enum { MGDT_RS_COPY = 0, MGDT_RS_AVGPOOL = 1, MGDT_RS_BILINEAR = 2, MGDT_RS_NEAREST = 3 };

int mgdt_abi_version(void);
const char* mgdt_last_error(void);
/* Number of kernels this library has enqueued so far in this process (every launch counts once;
 * a launch recorded into a CUDA graph counts when recorded, not per replay). */
unsigned long long mgdt_launch_count(void);
/* Debug: device buffer of 64 u64 per CTA receiving %globaltimer stamps of the tcgen05 conv's phases (NULL = off). */
void mgdt_debug_set_trace(void* buf);
/* Compiled-in facts for tests: returns 1 if the tcgen05/TMA conv path was built. */
int mgdt_has_umma(void);
/* Programmatic dependent launch (every kernel is launched with the programmatic-stream-serialization attribute and
 * orders itself with griddepcontrol.wait): 1 = on (default), 0 = plain launches. */
void mgdt_set_pdl(int on);
/* Library switches (A/B runs, debugging); the library itself reads no environment variables -- the Python layer
 * forwards MGDT_<NAME> once at load time.  Returns 0, or -EINVAL for an unknown name.
 *   "pdl"            1   programmatic dependent launch
 *   "conv_tma_load"  1   TMA-fed kernel (cp.async.bulk.tensor loads) for transform-free 1x1 convolutions
 *   "conv_tma_stats" 0   ... also for layers with fused output statistics (slower than the cp.async kernel's 16-warp epilogue)
 *   "conv_ksplit"    0   K = 16 steps of a tile rotate over 2 / 4 partial TMEM accumulators summed by the epilogue (an
 *                        experiment: measured no gain, tcgen05.mma accumulation chains are not the bound)
 *   "conv_tma_store" 1   TMA tensor stores of 1x1 epilogue units
 *   "conv_pair"      1   paired 16-column epilogue units (Cout <= 16)
 *   "conv_split"    -1   force the producer / epilogue warp split of the cp.async conv kernel (0 / 1 / 2)
 *   "conv3x3_warp"   2   3x3 stride-1 layers with Cin = Cout in {8, 16, 32} on the warp-level MMA kernel (0: tcgen05; 1: 8 / 16 only)
 *   "conv3x3_warp_spc" 1 strips per persistent CTA of the stride-1 warp kernel (> 1: double-buffered staging; no gain measured) */
int mgdt_set_option(const char* name, int value);

/* ---------------------------------------------------------------- convolution
 * Replaces Conv.forward/forward_fuse (nn/modules/conv.py:36-42) with BatchNorm folded
 * (yolo/utils/torch_utils.py:114-135), Bottleneck's residual add (nn/modules/block.py:524-526),
 * the raw nn.Conv2d leaves of Detect/TOODHead (nn/modules/head.py:150-151,484-493), the
 * nn.Linear layers of ConvNeXtV2_Block (nn/modules/convnextv2.py:27,30), MSPA_C2f's
 * `sp + spx[i]` (block.py:253), TaskDecomposition's per-sample reweighted 1x1 (head.py:122-126),
 * GRN folded into pwconv2 (nn/modules/utils.py:179-182) and `cls_feat * cls_prob` /
 * `F.relu(reg_feat)` in front of cv3/cv2 (head.py:528).
 *
 *   a(n,h,w,c) = relu?( (x + pre_add) * in_scale[n,c] * pix_scale[n,h,w] )
 *   y = act( conv(a, w) + bias ) + residual
 *
 * w is OHWI [Cout][kh][kw][Cin] in `dtype`; bias fp32 [Cout] or NULL; in_scale fp32 [N][Cin].
 */
typedef struct mgdt_conv_args {
    const void* x;         /* (N,H,W,Cin) channel stride x_cs */
    const void* w;
    const float* bias;
    void* y;               /* (N,Ho,Wo,Cout) channel stride y_cs */
    const void* pre_add;   /* NULL or same shape as x, stride add_cs */
    const float* in_scale; /* NULL or [N][Cin] */
    const void* pix_scale; /* NULL or (N,H,W,1) stride ps_cs */
    const void* residual;  /* NULL or same shape as y, stride res_cs */
    int32_t N, H, W, Cin, Cout;
    int32_t kh, kw, stride, pad;
    int32_t x_cs, y_cs, add_cs, ps_cs, res_cs;
    int32_t act, in_relu, dtype;
    int32_t impl;          /* 0 auto, 1 force CUDA-core path, 2 force tcgen05 path */
    const void* w_umma;    /* NULL, or the weights packed by mgdt_conv_umma_pack (bf16 tcgen05 path) */
    int32_t w_umma_f16;    /* 1 if w_umma was packed as fp16 (B operand F16, A stays bf16): 8x finer weight rounding */
    /* Fused output statistics (tcgen05 path only; mgdt_conv2d fails with -ENOTSUP otherwise, ask mgdt_conv2d_path first):
     * the epilogue adds, per image n and output channel c, the sums of the bf16-rounded outputs into
     * stat_acc[n][stat_q + stat_sq][Cout] (fp64, atomics): stat_q = 0 none, 1 total, 5 total + the four
     * adaptive_avg_pool2d(2) windows (q00, q01, q10, q11); stat_sq = 1 appends the sum of squares as the last plane.
     * The accumulators are replicated stat_copies (>= 1) times, stat_acc[copy][n][plane][c] with copy = tile % copies,
     * so that concurrent CTAs do not serialise on one L2 line; with stat_q = 5 and even Ho, Wo the total plane is left
     * untouched (the windows partition the image).  All copies must be zero on entry; mgdt_stats_finish sums them in
     * copy order, zeroes them again and runs the consumer (SPR gate / GRN scale / GroupNorm affine).  Replaces a
     * separate mgdt_chan_stats pass over y. */
    void* stat_acc;
    int32_t stat_q, stat_sq, stat_copies;
    /* Per-image weights (tcgen05 path, transform-free loader): w_umma points to N images packed back to back by
     * mgdt_conv_umma_pack_scaled (image n = the weights with input channel ci scaled by in_scale[n][ci]); in_scale,
     * pre_add, pix_scale and in_relu must then be unset.  The same product as in_scale on the activations
     * (W (s_n o x) = (W diag(s_n)) x) without the loader's in-place transform; tiles are cut per image. */
    int32_t w_per_image;
    /* act applies to output channels < act_cols only (0 = all); the rest get no activation.  Lets two 1x1 convs on the
     * same input run as one launch although only one of them is followed by an activation (InjectionMultiSum's
     * global_act -> h_sigmoid next to global_embedding, block.py:381-393).  Must be a multiple of 32; implemented by the
     * TMA-fed 1x1 kernel (mgdt_conv2d_path == 4), -ENOTSUP elsewhere. */
    int32_t act_cols;
} mgdt_conv_args;
int mgdt_conv2d(const mgdt_conv_args* a, void* stream);
/* Which kernel mgdt_conv2d would run for these arguments: 3 = conv_pointwise_kernel (narrow 1x1 layers, CUDA cores,
 * HBM-bound), 4 = conv1x1_tma_kernel (tcgen05, operand A fed by TMA: transform-free 1x1 layers), 5 = conv3x3_tma_kernel
 * (tcgen05, operand A fed by 4-D TMA boxes: transform-free 3x3 stride-1 layers with Cin <= 64), 2 = conv_umma2_kernel
 * (tcgen05, cp.async-fed: wide 3x3, stride 2, fused input transforms / statistics), 6 = conv3x3_warp_kernel (warp-level
 * mma.sync: transform-free 3x3 stride-1 layers with Cin = Cout in {8, 16, 32}, the Bottleneck pairs), 1 = conv_direct_kernel
 * (CUDA cores).  Used by the bench to attribute launches to kernels. */
int mgdt_conv2d_path(const mgdt_conv_args* a);

/* tcgen05 path: K-major shared-memory image of the weights, [col split][16-byte K chunk][Nc][8].
 * mgdt_conv_umma_packed_bytes returns 0 when (Cin, Cout, k, stride) is not taken by that path
 * (needs bf16, Cin % 8 == 0, k in {1,3} with pad k/2, stride 1 or (k=3) 2). */
size_t mgdt_conv_umma_packed_bytes(int Cin, int Cout, int k, int stride);
int mgdt_conv_umma_pack(const void* w_ohwi, int w_dtype, int Cin, int Cout, int k, int stride, int out_f16, void* packed,
                        void* stream);
/* N images of mgdt_conv_umma_packed_bytes each: image n = the bf16 image `packed_bf16` (from mgdt_conv_umma_pack) with
 * every input channel ci scaled by in_scale[n][ci] (fp32 [N][Cin]) and rounded to bf16 again. */
int mgdt_conv_umma_pack_scaled(const void* packed_bf16, int Cin, int Cout, int k, int stride, const float* in_scale, int N,
                               void* out, void* stream);
/* Same with the output columns scaled in groups of group_cols: columns [g*group_cols, (g+1)*group_cols), g < ngroups, use
 * in_scale[g][n][ci] (fp32 [ngroups][N][Cin]); columns past the last group stay unscaled.  Lets sibling 1x1 convs that
 * read the same map run as ONE GEMM with per-image weights -- TOODHead's cls_decomp / reg_decomp reduction convs (each
 * with its own layer attention) and cls_prob_conv1 (none), nn/modules/head.py:509-521. */
int mgdt_conv_umma_pack_scaled_groups(const void* packed_bf16, int Cin, int Cout, int k, int stride, const float* in_scale,
                                      int N, int ngroups, int group_cols, void* out, void* stream);

/* Fused input preprocessing + stem convolution on the tensor cores (bf16): 3x3 stride-2 pad-1 Conv+BN+act
 * (layer 0 of every config, models/v8/*.yaml) read straight from the NCHW uint8 (divided by 255,
 * predictor.py:127-129) or float32 source.  w_umma = mgdt_conv_umma_pack of the OHWI weights viewed as a
 * 1x1 conv over round_up(9*C, 16) channels (k = (dy*3+dx)*C + c, zero padded). */
int mgdt_stem_conv(const void* src, int src_is_u8, const void* w_umma, int w_umma_f16, const float* bias, void* y, int y_cs,
                   int N, int C, int H, int W, int Cout, int act, int dtype, void* stream);

/* The same layer for the predictor's own input -- uint8 NCHW, 3 channels (yolo/engine/predictor.py:115-130: /255 after
 * the copy) -- on warp-level tensor-core MMAs (csrc/stem_mma.cu): raw byte rows staged in shared memory, bytes turned
 * into exact fp16 integers in registers, fp16 weights scaled per output channel by a power of two, 1 / (255 * scale) on
 * the fp32 accumulator.  w = BN-folded fp32 weights (Cout, kp), k = (ky * 3 + kx) * 3 + c, kp >= 27 the row pitch.
 * Supported: C == 3, W % 16 == 0, Cout in {16, 32, 48, 64, 80}, y_cs % 8 == 0 (mgdt_stem_u8_supported); other shapes and
 * float32 sources take mgdt_stem_conv.  packed: mgdt_stem_u8_packed_bytes(Cout) bytes, 16-byte aligned. */
int mgdt_stem_u8_supported(int C, int H, int W, int Cout, int y_cs);
size_t mgdt_stem_u8_packed_bytes(int Cout);
int mgdt_stem_u8_pack(const float* w, int kp, int Cout, void* packed, void* stream);
int mgdt_stem_u8(const void* src, const void* packed, const float* bias, void* y, int y_cs, int N, int H, int W, int Cout,
                 int act, void* stream);

/* TOODHead classification tail in one launch (bf16): logits = cv3(cls_feat * sigmoid(cls_prob_conv2(prob))), nn/modules/
 * head.py:519-521, 528.  prob (N,H,W,C1) is cls_prob_conv1's ReLU output, w2 the OHWI (1,3,3,C1) weights of
 * cls_prob_conv2 (3x3, pad 1), feat (N,H,W,C2) the classification feature, w3 the (nc,C2) weights of cv3; out receives nc
 * channels per pixel (a slice of the head's raw map, channel stride out_cs).  C1, C2 multiples of 8 (<= 64 / 256),
 * nc <= 8 (mgdt_tood_cls_supported); otherwise run the two convolutions (mgdt_conv2d with pix_scale). */
int mgdt_tood_cls_supported(int C1, int C2, int nc, int prob_cs, int feat_cs);
int mgdt_tood_cls(const void* prob, int prob_cs, const void* w2, const float* b2, const void* feat, int feat_cs,
                  const void* w3, const float* b3, void* out, int out_cs, int N, int H, int W, int C1, int C2, int nc,
                  int dtype, void* stream);

/* MSPA_C2f hierarchy front (nn/modules/block.py:248-262) in ONE launch (bf16): the chain of pointwise Conv+BN+act
 * branches  sp_0 = convs[0](spx[0]);  sp_i = convs[i](sp_{i-1} + spx[i]), i < nstage;  sp_in = sp_{nstage-1} + spx[nstage]
 * where spx[i] = x[:, i*iw:(i+1)*iw].  x has (nstage+1)*iw channels; sp_i is written to ycat[:, i*iw:(i+1)*iw]
 * (the concat buffer convs[-1] reads), sp_in (the bottleneck input) to ysp (iw channels).  A warp owns 32 pixels and
 * chains warp-level bf16 MMAs in registers (the accumulator fragment of one stage is the A fragment of the next).
 * w_packed = mgdt_mspa_front_pack of the fp32 [nstage][ci][co] weights (BN folded): bf16 B fragments; bias fp32
 * [nstage][iw].  Intermediate values are rounded to bf16 where the unfused sequence (mgdt_conv2d with pre_add,
 * mgdt_affine_act) stores / stages them.  iw in {8, 16, 32, 64}. */
int mgdt_mspa_front_supported(int iw, int nstage);
size_t mgdt_mspa_front_packed_bytes(int iw, int nstage);
int mgdt_mspa_front_pack(const float* w, int nstage, int iw, void* packed, void* stream);
int mgdt_mspa_front(const void* x, int x_cs, const void* w_packed, const float* bias, int nstage, int iw, int act, void* ycat,
                    int y_cs, void* ysp, int s_cs, int N, int H, int W, int dtype, void* stream);

/* Depthwise 7x7 (pad 3, bias) + channels-last LayerNorm(eps), ConvNeXtV2_Block.forward
 * (nn/modules/convnextv2.py:35-37, nn/modules/utils.py:162-163).  w is [49][C] in dtype,
 * bias/ln_w/ln_b fp32 [C]. */
int mgdt_dwconv7_ln(const void* x, int x_cs, const void* w, const float* bias, const float* ln_w, const float* ln_b,
                    float eps, void* y, int y_cs, int N, int H, int W, int C, int dtype, void* stream);

/* Modulated deformable 3x3 conv, stride 1, pad 1 (DyDCNv2.forward, nn/modules/block.py:427-429;
 * arithmetic of mmcv ModulatedDeformConv2d / torchvision.ops.deform_conv2d).  offset: 18 channels
 * ((dy,dx) per tap), mask: 9 channels; with mask_is_logit the sigmoid of head.py:517 is applied
 * here, so both may be channel slices of the raw spatial_conv_offset output (head.py:515-517).
 * w is [Cout][9][Cin] in dtype; w_umma (optional, bf16) is that matrix packed by mgdt_conv_umma_pack as
 * a 1x1 conv over 9*Cin channels: the tensor-core path builds the modulated bilinear im2col tile in
 * shared memory and runs it through tcgen05.  stat_acc / stat_q / stat_sq / stat_copies: fused output statistics as
 * in mgdt_conv_args (NULL = none; the GroupNorm that follows DyDCNv2, block.py:430). */
int mgdt_dcn3x3(const void* x, int x_cs, const void* offset, int off_cs, const void* mask, int mask_cs,
                int mask_is_logit, const void* w, const void* w_umma, int w_umma_f16, void* y, int y_cs, int N, int H, int W,
                int Cin, int Cout, int dtype, void* stat_acc, int stat_q, int stat_sq, int stat_copies, void* stream);
/* 2 if mgdt_dcn3x3 would run the tcgen05 path for these arguments (the only one that takes stat_acc), else 1. */
int mgdt_dcn3x3_path(const void* x, int x_cs, const void* w_umma, int N, int H, int W, int Cin, int Cout, int dtype);

/* ---------------------------------------------------------------- reductions
 * Per-(n,c) sums over the image, optionally per adaptive 2x2 window as well.
 * quads=0: out_sum[N][1][C];  quads=1: out_sum[N][5][C] = {total, q00, q01, q10, q11} with
 * adaptive_avg_pool2d(2) windows [floor(i*H/2), ceil((i+1)*H/2)).  out_sumsq (same layout, total
 * only -> [N][C]) may be NULL.  Feeds SPRModule (spr_module.py:22-24), GRN (utils.py:180),
 * GroupNorm (head.py:76, block.py:425) and TaskDecomposition's GAP (head.py:507).
 * One launch, deterministic: blocks write chunk partials into `ws` (mgdt_chan_stats_ws_bytes()), the last block of
 * each image (atomic ticket) sums them in chunk order.  counters: int32[N] tickets, ZERO on entry and left zero on
 * exit; an array may be reused by consecutive calls on one stream, not by calls that can run concurrently. */
size_t mgdt_chan_stats_ws_bytes(int N, int H, int W, int C, int quads);
int mgdt_chan_stats(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum, float* out_sumsq,
                    void* ws, size_t ws_bytes, int32_t* counters, int dtype, void* stream);

/* chan_stats whose last block also runs the per-image computation that consumes the statistics (one launch saved):
 *   MGDT_FIN_GATE  SPRModule MLP + softmax over groups (== mgdt_mspa_gate): p0..p3 = fc1 w, fc1 b, fc2 w, fc2 b,
 *                  i0 = groups, i1 = softmax, i2 = hidden, o0 = scale[N][C]; needs quads = 1
 *   MGDT_FIN_GRN   GRN scale (== mgdt_grn_scale): p0 = gamma, o0 = scale[N][C]; needs out_sumsq
 *   MGDT_FIN_GN    GroupNorm affine (== mgdt_gn_affine): p0 = gamma, p1 = beta, i0 = groups, f0 = eps,
 *                  o0 = a[N][C], o1 = b[N][C]; needs out_sumsq, quads = 0 */
enum { MGDT_FIN_NONE = 0, MGDT_FIN_GATE = 1, MGDT_FIN_GRN = 2, MGDT_FIN_GN = 3 };
typedef struct mgdt_stats_fin {
    int32_t kind;
    const float *p0, *p1, *p2, *p3;
    int32_t i0, i1, i2;
    float f0;
    float *o0, *o1;
} mgdt_stats_fin;
int mgdt_chan_stats_fin(const void* x, int x_cs, int N, int H, int W, int C, int quads, float* out_sum, float* out_sumsq,
                        void* ws, size_t ws_bytes, int32_t* counters, const mgdt_stats_fin* fin, int dtype, void* stream);

/* Consumer of the statistics a convolution accumulated in its epilogue (mgdt_conv_args.stat_acc): reduces
 * acc[copies][N][q + sq][C] (fp64) to out_sum[N][q][C] / out_sumsq[N][C] (either may be NULL when its planes are
 * absent; H, W are the map's size: for q = 5 and even H, W the total is the sum of the four windows), zeroes acc for
 * the next use, then runs the finaliser `fin` (same kinds and arguments as mgdt_chan_stats_fin; NULL or kind 0 =
 * none).  One block per image. */
int mgdt_stats_finish(void* acc, int copies, int N, int H, int W, int C, int q, int sq, float* out_sum, float* out_sumsq,
                      const mgdt_stats_fin* fin, void* stream);

/* SPR gate of MSPA_C2f (block.py:270-279 + spr_module.py:20-31): stats[N][5][C] (sums) ->
 * scale[N][C] = softmax over the `groups` (4) channel groups of sigmoid(fc2(relu(fc1([mean | 2x2 means])))).
 * fc1_w [hidden][5*ow], fc2_w [ow][hidden] fp32, ow = C/groups.  groups=1, softmax=0 gives
 * SPRModule.forward alone. */
int mgdt_mspa_gate(const float* stats, int N, int H, int W, int C, int groups, int softmax, const float* fc1_w,
                   const float* fc1_b, const float* fc2_w, const float* fc2_b, int hidden, float* scale, void* stream);

/* GRN (utils.py:179-182) as a per-(n,c) input scale for pwconv2: s = 1 + gamma*Gx/(mean_c Gx + 1e-6),
 * Gx = sqrt(sumsq).  (The beta term is folded into pwconv2's bias by the host.) */
int mgdt_grn_scale(const float* sumsq, const float* gamma, int N, int C, float* scale, void* stream);

/* GroupNorm finalize: per-(n,c) affine a,b with y = x*a + b  (nn.GroupNorm(groups, C), eps). */
int mgdt_gn_affine(const float* sum, const float* sumsq, int N, int C, int groups, int hw, float eps,
                   const float* gamma, const float* beta, float* a, float* b, void* stream);

/* TaskDecomposition layer attention (head.py:116-117): sum[N][C] -> in_scale[which][N][C] for both
 * `ndec` decompositions at once (weights stacked): sigmoid(la2(relu(la1(mean))))[c / (C/stacked)]. */
int mgdt_td_attn(const float* sum, int N, int C, int hw, int hidden, int stacked, int ndec, const float* la1_w,
                 const float* la1_b, const float* la2_w, const float* la2_b, float* in_scale, void* stream);

/* ---------------------------------------------------------------- elementwise / gather
 * y = act(x * a[n,c] + b[n,c]) (+ other);  a/b may be NULL (1 / 0).  Used for GroupNorm apply
 * (+SiLU / +ReLU), the MSPA attention scale (block.py:279) and MSPA's sp + spx[3]. */
int mgdt_affine_act(const void* x, int x_cs, const float* a, const float* b, const void* other, int o_cs, int act,
                    void* y, int y_cs, int N, int H, int W, int C, int dtype, void* stream);

/* Resample x (N,Hi,Wi,C) into y (N,Ho,Wo,C): copy / adaptive average pool / bilinear
 * (align_corners=False) / nearest.  SimFusion_4in/3in (block.py:294-329), nn.Upsample + Concat of
 * the PAN neck (models/v8/yolov8.yaml:30-46), Concat (conv.py:287-297). */
int mgdt_resample(const void* x, int x_cs, int Hi, int Wi, void* y, int y_cs, int Ho, int Wo, int N, int C, int mode,
                  int dtype, void* stream);

/* SPPF pooling chain (block.py:151-153): y1 = maxpool5(x), y2 = maxpool5(y1), y3 = maxpool5(y2),
 * i.e. 5x5 / 9x9 / 13x13 windows (k=5).  x and y1..y3 are slices with the given strides. */
int mgdt_sppf_pool(const void* x, int x_cs, void* y1, void* y2, void* y3, int y_cs, int N, int H, int W, int C, int k,
                   int dtype, void* stream);

/* InjectionMultiSum_Auto_pool tail (block.py:385-396): out = local * G(act) + G(feat) where
 * G = bilinear(align_corners=False) of (h_sigmoid(act), feat) when Hg <= H, else adaptive avg pool
 * of (act, feat) without h_sigmoid. */
int mgdt_inject(const void* local, int l_cs, const void* gact, int a_cs, const void* gfeat, int f_cs, void* y, int y_cs,
                int N, int H, int W, int Hg, int Wg, int C, int dtype, void* stream);
/* Same; gact_is_hsig = 1: `gact` already holds h_sigmoid(global_act), applied by the producing convolution's epilogue
 * (mgdt_conv_args.act = HSIGMOID with act_cols): bf16 exact-2x upsampling only, -ENOTSUP otherwise. */
int mgdt_inject2(const void* local, int l_cs, const void* gact, int a_cs, const void* gfeat, int f_cs, void* y, int y_cs,
                 int N, int H, int W, int Hg, int Wg, int C, int gact_is_hsig, int dtype, void* stream);

/* uint8 NCHW -> NHWC float/bf16, scaled by 1/255 (BasePredictor.preprocess,
 * yolo/engine/predictor.py:115-130).  Also float NCHW -> NHWC (scale 1). */
int mgdt_preprocess(const void* src, int src_is_u8, void* y, int y_cs, int N, int C, int H, int W, int dtype,
                    void* stream);

/* ---------------------------------------------------------------- decode
 * Detect/TOODHead inference tail (head.py:165-177,536-559) = DFL (block.py:50-53) + make_anchors /
 * dist2bbox(xywh) (yolo/utils/tal.py:476-500) + sigmoid + cat.  y is fp32 (N, 4+nc, A).
 * dist_only=1: DFL.forward alone, y is (N, 4, A) ltrb expectations. */
typedef struct mgdt_decode_level {
    const void* raw; /* (N,H,W,4*reg_max+nc) */
    int32_t H, W, cs;
    float stride;
} mgdt_decode_level;
int mgdt_decode(const mgdt_decode_level* levels, int nl, int N, int reg_max, int nc, int dist_only, float* y, int dtype,
                void* stream);

/* ---------------------------------------------------------------- NMS
 * non_max_suppression (yolo/utils/ops.py:136-266) + torchvision.ops.nms, whole batch, no host
 * sync.  pred fp32 (N, 4+nc, A).  out fp32 (N, max_det, 6) rows (x1,y1,x2,y2,conf,cls); counts
 * int32 [N].  classes: NULL or int32 list of allowed class ids.  Score ties are broken by
 * candidate index (anchor-major, then class), i.e. a stable sort. */
size_t mgdt_nms_ws_bytes(int N, int nc, int A, int multi_label, int max_nms);
int mgdt_nms(const float* pred, int N, int nc, int A, float conf_thres, float iou_thres, int multi_label, int agnostic,
             int max_det, int max_nms, float max_wh, const int32_t* classes, int n_classes, float* out,
             int32_t* counts, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- pre / post-processing
 * LetterBox (yolo/data/augment.py:538-593: cv2.resize INTER_LINEAR to (new_h, new_w), constant border) fused with the
 * BGR->RGB / HWC->CHW of BasePredictor.preprocess (yolo/engine/predictor.py:121-125).  src is one HWC uint8 image with
 * 3 channels and `pitch` bytes per row; dst is one (3, H, W) image slot of the NCHW batch (or an (H, W, 3) image when
 * out_hwc).  The geometry (new size, top / left border) is computed by the caller exactly as LetterBox does.
 * Bit-exact with OpenCV's fixed-point uint8 bilinear resize. */
int mgdt_letterbox_u8(const void* src, int h0, int w0, int pitch, void* dst, int H, int W, int new_h, int new_w, int top,
                      int left, int swap_rb, int pad_value, int out_hwc, void* stream);

/* ops.scale_boxes + clip_boxes (yolo/utils/ops.py:90-117, 269-285) in place on packed detections
 * dets[N][max_rows][row_stride] (xyxy first); counts[N] rows are valid per image (NULL = all max_rows);
 * params[N][5] = gain, pad_w, pad_h, h0, w0 as the reference computes them. */
int mgdt_scale_boxes(float* dets, int row_stride, const int32_t* counts, int N, int max_rows, const float* params,
                     void* stream);

/* Boxes.xywh / xyxyn / xywhn (yolo/engine/results.py:405-430; ops.xyxy2xywh, yolo/utils/ops.py:345-359): converts the
 * xyxy columns of n rows of `boxes` (row_stride floats per row) into out[n][4].  mode bit 0: xyxy -> xywh (centre,
 * size); bit 1: x / w, y / h.  Bit-exact with the reference's fp32 arithmetic. */
int mgdt_box_convert(const float* boxes, int row_stride, int n, int mode, float w, float h, float* out, void* stream);

/* DetectionValidator._process_batch (yolo/v8/detect/val.py:150-175) with metrics.box_iou (yolo/utils/metrics.py:52-72)
 * for a whole batch in one launch.  dets[N][max_det][det_stride] rows (x1,y1,x2,y2,conf,cls,...) in native image space
 * (after mgdt_scale_boxes), det_counts[N] valid rows (NULL = max_det); labels[N][max_lab][5] rows (cls,x1,y1,x2,y2) in
 * the same space, lab_counts[N] (NULL = max_lab); iouv[niou] the IoU levels (torch.linspace(0.5, 0.95, 10), val.py:28).
 * correct[N][max_det][niou] (uint8 0/1; rows past the count are 0): per level every detection keeps its highest-IoU
 * class-matching label with IoU >= level, every label then keeps the lowest-index detection that chose it.  IoU ties
 * between two labels of one detection keep the lower label index (the reference's argsort leaves them unspecified). */
int mgdt_match_batch(const float* dets, int det_stride, const int32_t* det_counts, int max_det, const float* labels,
                     const int32_t* lab_counts, int max_lab, const float* iouv, int niou, uint8_t* correct, int N,
                     void* stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * Row f3: training criterion and optimizer-side element-wise work (csrc/train.cu).
 *
 * mgdt_v8_loss replaces v8DetectionLoss.__call__ (yolo/utils/loss.py:159-208) after the concatenation of the head's
 * train-mode outputs: bbox_decode (loss.py:150-157), HeuristicPositiveSampleAssigner_v1 / TaskAlignedAssigner
 * (yolo/utils/tal.py:56-142, 144-353; CIoU overlaps of yolo/utils/metrics.py:75-128, top-k, select_highest_overlaps on
 * the alignment metric, normalised target scores), BCE + CIoU + DFL (loss.py:60-92) -- and their gradient.
 *   pred      (B, no, A) fp32, no = 4 * reg_max + nc: the DFL logits (side-major) then the class logits
 *   anchors   (A, 2) fp32 anchor centres in grid units, strides (A) fp32   (make_anchors, tal.py:484-500)
 *   gt        (B, G, 5) fp32 rows (class, x1, y1, x2, y2) in pixels, all-zero rows = padding (loss.py:131-148); G may be 0
 *   alpha     0.5 * (100 - calls / 161) / 100 (tal.py:110, 266), beta 8, topk 10 as the reference constructs it
 *   loss4     out: box, cls, dfl (each already multiplied by its gain; the reference returns their sum * B) and
 *             target_scores_sum
 *   grad_pred out or NULL: d(sum(loss3) * B) / d(pred), same layout as pred
 *   out_tscore / out_idx / out_tbox / out_label: optional assigner outputs, (B, A) fp32 target score of the assigned
 *             class, (B, A) int32 gt index or -1, (B, A, 4) fp32 target box in pixels, (B, A) int32 label or -1
 *   ws        device scratch of mgdt_v8_loss_ws_bytes(B, A, G) bytes */
size_t mgdt_v8_loss_ws_bytes(int B, int A, int G);
int mgdt_v8_loss(const float* pred, const float* anchors, const float* strides, const float* gt, int B, int A, int G, int nc,
                 int reg_max, float alpha, float beta, int topk, float box_gain, float cls_gain, float dfl_gain, float* loss4,
                 float* grad_pred, float* out_tscore, int32_t* out_idx, float* out_tbox, int32_t* out_label, void* ws, size_t ws_bytes,
                 void* stream);
/* ModelEMA.update (yolo/utils/torch_utils.py:347-358) over one flat fp32 buffer: ema = ema * decay + (1 - decay) * model. */
int mgdt_ema_update(float* ema, const float* model, size_t n, float decay, void* stream);
/* Sum of squares of a flat fp32 buffer into *out (fp64; zeroed first): the gradient norm of clip_grad_norm_. */
int mgdt_sumsq(const float* x, size_t n, double* out, void* stream);
/* DetectionTrainer.optimizer_step (yolo/engine/trainer.py:462-470) over one flat bucket: clip_grad_norm_(max_norm) from
 * *gnorm_sq (NULL = no clipping), then torch.optim.SGD(momentum, nesterov) with the three parameter groups of
 * build_optimizer (trainer.py:614-650): group[i] = 0 weights (decay), 1 normalisation weights, 2 biases; lr3 / wd3 are
 * HOST arrays of three floats.  grad_scale multiplies the gradients first (1 / world_size after a summing all-reduce). */
int mgdt_sgd_step(float* prm, const float* grad, float* mom, const unsigned char* group, size_t n, const float* lr3, const float* wd3,
                  float momentum, int nesterov, int first_step, const double* gnorm_sq, float max_norm, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGDT_B200_H */
