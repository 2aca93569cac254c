// C-ABI glue: error strings, version, conv dispatch between the CUDA-core and tcgen05 paths.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

namespace mgdt {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = 1;

int pdl_enabled() { return g_pdl != 0; }   // on unless mgdt_set_pdl(0) / mgdt_set_option("pdl", 0)

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int blocks_set_option(const char* name, int value);
int conv2d_direct(const mgdt_conv_args* a, cudaStream_t s);
bool conv2d_pointwise_supported(const mgdt_conv_args* a);
int conv2d_pointwise(const mgdt_conv_args* a, cudaStream_t s);
extern int g_conv3x3_warp, g_conv3x3_warp_spc;
bool conv3x3_warp_supported(const mgdt_conv_args* a);
int conv3x3_warp(const mgdt_conv_args* a, cudaStream_t s);
#ifdef MGDT_WITH_UMMA
int conv_set_option(const char* name, int value);
bool conv2d_umma_supported(const mgdt_conv_args* a);
int conv2d_umma(const mgdt_conv_args* a, cudaStream_t s);
int conv2d_umma_path(const mgdt_conv_args* a);
#endif

}  // namespace mgdt

using namespace mgdt;

extern "C" int mgdt_abi_version(void) { return MGDT_ABI_VERSION; }
extern "C" const char* mgdt_last_error(void) { return g_err; }
extern "C" unsigned long long mgdt_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" void mgdt_set_pdl(int on) { g_pdl = on ? 1 : 0; }
extern "C" int mgdt_set_option(const char* name, int value) {
    MGDT_CHECK(name, "set_option: null name");
    if (!strcmp(name, "pdl")) { g_pdl = value ? 1 : 0; return 0; }
    if (!strcmp(name, "conv3x3_warp")) { g_conv3x3_warp = value; return 0; }
    if (!strcmp(name, "conv3x3_warp_spc")) { g_conv3x3_warp_spc = value; return 0; }
    if (blocks_set_option(name, value)) return 0;
#ifdef MGDT_WITH_UMMA
    if (conv_set_option(name, value)) return 0;
#endif
    return set_error(-EINVAL, "set_option: unknown option '%s'", name);
}
extern "C" int mgdt_conv2d_path(const mgdt_conv_args* a) {
    if (a && a->impl == 0 && !a->stat_acc && !a->w_per_image && !a->act_cols && conv2d_pointwise_supported(a)) return 3;
    if (a && conv3x3_warp_supported(a)) return 6;
#ifdef MGDT_WITH_UMMA
    if (a && a->impl != 1 && conv2d_umma_supported(a)) return conv2d_umma_path(a);
#endif
    return 1;
}
extern "C" int mgdt_has_umma(void) {
#ifdef MGDT_WITH_UMMA
    return 1;
#else
    return 0;
#endif
}

extern "C" int mgdt_conv2d(const mgdt_conv_args* a, void* stream) {
    MGDT_CHECK(a, "conv2d: null args");
    MGDT_CHECK(a->x && a->w && a->y, "conv2d: null tensor pointer");
    MGDT_CHECK(a->N > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, "conv2d: bad shape");
    MGDT_CHECK(a->kh > 0 && a->kw > 0 && a->stride > 0 && a->pad >= 0, "conv2d: bad kernel/stride/pad");
    MGDT_CHECK(a->H + 2 * a->pad >= a->kh && a->W + 2 * a->pad >= a->kw, "conv2d: kernel larger than padded input");
    MGDT_CHECK(a->x_cs >= a->Cin && a->y_cs >= a->Cout, "conv2d: channel stride smaller than channel count");
    MGDT_CHECK(!a->pre_add || a->add_cs >= a->Cin, "conv2d: bad pre_add stride");
    MGDT_CHECK(!a->residual || a->res_cs >= a->Cout, "conv2d: bad residual stride");
    MGDT_CHECK(!a->pix_scale || a->ps_cs >= 1, "conv2d: bad pix_scale stride");
    MGDT_CHECK(a->act >= MGDT_ACT_NONE && a->act <= MGDT_ACT_GELU, "conv2d: bad act %d", a->act);
    cudaStream_t s = (cudaStream_t)stream;
    MGDT_CHECK(!a->stat_acc || ((a->stat_q == 0 || a->stat_q == 1 || a->stat_q == 5) && a->stat_q + (a->stat_sq ? 1 : 0) > 0 &&
                                ((uintptr_t)a->stat_acc & 7) == 0), "conv2d: bad fused-statistics request");
    if (a->impl == 0 && !a->stat_acc && !a->w_per_image && !a->act_cols && conv2d_pointwise_supported(a)) return conv2d_pointwise(a, s);   // narrow 1x1 layers: HBM-bound SIMT
    if (conv3x3_warp_supported(a)) return conv3x3_warp(a, s);   // 3x3 s1 with 8 / 16 / 32 channels: warp-level MMAs, no per-launch set-up
#ifdef MGDT_WITH_UMMA
    if (a->impl != 1 && conv2d_umma_supported(a)) return conv2d_umma(a, s);
#endif
    if (a->stat_acc) return set_error(-ENOTSUP, "conv2d: fused statistics need the tcgen05 path (see mgdt_conv2d_path)");
    if (a->w_per_image) return set_error(-ENOTSUP, "conv2d: per-image weights need the tcgen05 path (see mgdt_conv2d_path)");
    if (a->act_cols) return set_error(-ENOTSUP, "conv2d: act_cols needs the TMA-fed 1x1 kernel (see mgdt_conv2d_path)");
#ifdef MGDT_WITH_UMMA
    MGDT_CHECK(a->impl != 2, "conv2d: tcgen05 path does not support this shape");
#else
    MGDT_CHECK(a->impl != 2, "conv2d: library built without the tcgen05 path");
#endif
    return conv2d_direct(a, s);
}
