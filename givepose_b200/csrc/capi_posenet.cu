// givepose_b200 -- C-ABI entry points of the PoseNet glue kernels (include/givepose_b200.h, "PoseNet forward" block).
#include "posenet_kernels.cuh"

using namespace gp;

static bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// slab decomposition shared by the channel-last elementwise kernels: grid (slabs, N), >= 32 pixels per CTA
static dim3 slab_grid(int N, int npix, int ctas_per_sm, int *ppc) {
    int slabs = (int)((148ll * ctas_per_sm + N - 1) / N);
    if (slabs > (npix + 31) / 32) slabs = (npix + 31) / 32;
    if (slabs < 1) slabs = 1;
    *ppc = (npix + slabs - 1) / slabs;
    return dim3((npix + *ppc - 1) / *ppc, N);
}

// statistics passes shared by the two GroupNorm entry points: partials -> (mean, rstd) in stats[0 .. N*G*2)
template <typename T>
static void gn_statistics(const void *x, float *stats, int N, int HW, int C, int G, float eps, cudaStream_t st) {
    int ppc = 0;
    const dim3 sgrid = slab_grid(N, HW, 8, &ppc);
    float *partial = stats + (size_t)N * G * 2;
    gn_stats_kernel<T><<<sgrid, 256, 256 * 2 * sizeof(float), st>>>((const T *)x, partial, HW, C, G, ppc);
    gn_finalize_kernel<<<(N * G + 127) / 128, 128, 0, st>>>(partial, stats, N * G, G, (int)sgrid.x, 1.f / ((float)HW * (C / G)), eps);
    count_launch(2);
}

extern "C" {

size_t gp_groupnorm_workspace_floats(int N, int H, int W, int G) {
    if (N <= 0 || H <= 0 || W <= 0 || G <= 0) return 0;
    int ppc = 0;
    const dim3 sgrid = slab_grid(N, H * W, 8, &ppc);
    return (size_t)N * G * 2 * (1 + sgrid.x);   // (mean, rstd) per (n, g) followed by the slab partials
}


int gp_dwconv3x3_ln_gelu(const void *x, const float *w_t, const float *bias, const float *ln_w, const float *ln_b, void *out,
                         int N, int H, int W, int C, long long rows, float eps, int dtype, void *stream) {
    if (!x || !w_t || !bias || !ln_w || !ln_b || !out) return GP_ERR_NULL;
    if (N <= 0 || H <= 0 || W <= 0 || rows < 0 || rows > (long long)N * H * W) return GP_ERR_SHAPE;
    if (C != 128 && C != 256 && C != 512) return GP_ERR_UNSUPPORTED;
    if (!al16(x) || !al16(w_t) || !al16(bias) || !al16(ln_w) || !al16(ln_b) || !al16(out)) return GP_ERR_ALIGN;
    if (rows == 0) return GP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long want = (rows + 7) / 8;   // 8 warps (pixels) per CTA
    const unsigned grid = (unsigned)(want < 148ll * 64 ? want : 148ll * 64);
#define GP_DW(TT, JJ) dwconv3x3_ln_gelu_kernel<TT, JJ><<<grid, 256, 0, st>>>((const TT *)x, w_t, bias, ln_w, ln_b, (TT *)out, H, W, rows, eps)
#define GP_DW_T(TT) do { if (C == 128) GP_DW(TT, 1); else if (C == 256) GP_DW(TT, 2); else GP_DW(TT, 4); } while (0)
    switch (dtype) {
        case GP_F32: GP_DW_T(float); break;
        case GP_BF16: GP_DW_T(__nv_bfloat16); break;
        case GP_F16: GP_DW_T(__half); break;
        default: return GP_ERR_DTYPE;
    }
#undef GP_DW_T
#undef GP_DW
    count_launch();
    return (int)cudaGetLastError();
}

int gp_small_k_linear(const void *x, const float *w_t, const float *bias, void *out, long long rows, int K, int C, int dtype,
                      void *stream) {
    if (!x || !w_t || !bias || !out) return GP_ERR_NULL;
    if (rows < 0 || C <= 0 || C % 8 || C / 4 > 256) return GP_ERR_SHAPE;
    if (K != 3) return GP_ERR_UNSUPPORTED;
    if (!al16(out)) return GP_ERR_ALIGN;
    if (rows == 0) return GP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int V = dtype == GP_F32 ? 4 : 8, pstep = 256 / (C / V);
    if (pstep < 1) return GP_ERR_SHAPE;
    const long long want = (rows + pstep - 1) / pstep;
    const unsigned grid = (unsigned)(want < 148ll * 32 ? want : 148ll * 32);
    switch (dtype) {
        case GP_F32: small_k_linear_kernel<float, 3><<<grid, 256, 0, st>>>((const float *)x, w_t, bias, (float *)out, rows, C); break;
        case GP_BF16: small_k_linear_kernel<__nv_bfloat16, 3><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, w_t, bias, (__nv_bfloat16 *)out, rows, C); break;
        case GP_F16: small_k_linear_kernel<__half, 3><<<grid, 256, 0, st>>>((const __half *)x, w_t, bias, (__half *)out, rows, C); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_smallk_dwconv3x3_ln_gelu(const void *x, const float *w_eff, const float *bias, const float *ln_w, const float *ln_b,
                                void *out, int N, int H, int W, int K, int C, long long rows, float eps, int dtype, void *stream) {
    if (!x || !w_eff || !bias || !ln_w || !ln_b || !out) return GP_ERR_NULL;
    if (N <= 0 || H <= 0 || W <= 0 || rows < 0 || rows > (long long)N * H * W) return GP_ERR_SHAPE;
    if (K != 3 || (C != 128 && C != 256 && C != 512)) return GP_ERR_UNSUPPORTED;
    if (!al16(w_eff) || !al16(bias) || !al16(ln_w) || !al16(ln_b) || !al16(out)) return GP_ERR_ALIGN;
    if (rows == 0) return GP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int px = (W % 4 == 0 && C <= 256) ? 4 : 1;   // pixels per warp: the composed weights are loaded once per 4 pixels
    const long long want = (rows + 8 * px - 1) / (8 * px);
    const unsigned grid = (unsigned)(want < 148ll * 64 ? want : 148ll * 64);
#define GP_SK(TT, JJ) do { if (px == 4) smallk_dwconv_ln_gelu_kernel<TT, JJ, 3, 4><<<grid, 256, 0, st>>>((const TT *)x, w_eff, bias, ln_w, ln_b, (TT *)out, H, W, rows, eps); \
                           else smallk_dwconv_ln_gelu_kernel<TT, JJ, 3, 1><<<grid, 256, 0, st>>>((const TT *)x, w_eff, bias, ln_w, ln_b, (TT *)out, H, W, rows, eps); } while (0)
#define GP_SK_T(TT) do { if (C == 128) GP_SK(TT, 1); else if (C == 256) GP_SK(TT, 2); else smallk_dwconv_ln_gelu_kernel<TT, 4, 3, 1><<<grid, 256, 0, st>>>((const TT *)x, w_eff, bias, ln_w, ln_b, (TT *)out, H, W, rows, eps); } while (0)
    switch (dtype) {
        case GP_F32: GP_SK_T(float); break;
        case GP_BF16: GP_SK_T(__nv_bfloat16); break;
        case GP_F16: GP_SK_T(__half); break;
        default: return GP_ERR_DTYPE;
    }
#undef GP_SK_T
#undef GP_SK
    count_launch();
    return (int)cudaGetLastError();
}

static int groupnorm_act_impl(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta, int N, int H,
                              int W, int C, int G, float eps, int act, int dtype, void *stream, bool have_stats) {
    if (!x || !y || !stats || !gamma || !beta) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || C <= 0 || G <= 0 || C % G || (C / G) % 4 || C / 4 > 256) return GP_ERR_SHAPE;
    if (act < 0 || act > 2) return GP_ERR_UNSUPPORTED;
    if (dtype != GP_F32 && C % 8) return GP_ERR_SHAPE;   // 16-byte channel vectors
    if (!al16(x) || !al16(y) || !al16(gamma) || !al16(beta)) return GP_ERR_ALIGN;
    if (stats_floats < (have_stats ? (size_t)N * G * 2 : gp_groupnorm_workspace_floats(N, H, W, G))) return GP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    int appc = 0;
    const dim3 agrid = slab_grid(N, HW, 16, &appc);
#define GP_GN_APPLY(TT, AA) gn_apply_kernel<TT, AA><<<agrid, 256, 0, st>>>((const TT *)x, stats, gamma, beta, (TT *)y, HW, C, G, eps, appc)
#define GP_GN(TT) do { if (!have_stats) gn_statistics<TT>(x, stats, N, HW, C, G, eps, st); \
                       if (act == ACT_RELU) GP_GN_APPLY(TT, ACT_RELU); else if (act == ACT_GELU) GP_GN_APPLY(TT, ACT_GELU); \
                       else GP_GN_APPLY(TT, ACT_NONE); } while (0)
    switch (dtype) {
        case GP_F32: GP_GN(float); break;
        case GP_BF16: GP_GN(__nv_bfloat16); break;
        case GP_F16: GP_GN(__half); break;
        default: return GP_ERR_DTYPE;
    }
#undef GP_GN
#undef GP_GN_APPLY
    count_launch();
    return (int)cudaGetLastError();
}

int gp_groupnorm_act(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta, int N, int H,
                     int W, int C, int G, float eps, int act, int dtype, void *stream) {
    return groupnorm_act_impl(x, y, stats, stats_floats, gamma, beta, N, H, W, C, G, eps, act, dtype, stream, false);
}

// apply pass only: stats[0 .. N*G*2) already holds (mean, rstd) per (n, group), e.g. from gp_conv3x3_gn_bf16 + gp_groupnorm_finalize
int gp_groupnorm_apply(const void *x, void *y, const float *stats, size_t stats_floats, const float *gamma, const float *beta, int N, int H,
                       int W, int C, int G, float eps, int act, int dtype, void *stream) {
    return groupnorm_act_impl(x, y, const_cast<float *>(stats), stats_floats, gamma, beta, N, H, W, C, G, eps, act, dtype, stream, true);
}

// (mean, rstd) per (n, group) from per-slab (sum, sum of squares) partials laid out [N][slabs][G][2], summed in slab order
int gp_groupnorm_finalize(const float *partial, float *stats, int N, int G, int slabs, long long count, float eps, void *stream) {
    if (!partial || !stats) return GP_ERR_NULL;
    if (N <= 0 || G <= 0 || slabs <= 0 || count <= 0) return GP_ERR_SHAPE;
    gn_finalize_kernel<<<(N * G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partial, stats, N * G, G, slabs, 1.f / (float)count, eps);
    count_launch();
    return (int)cudaGetLastError();
}

int gp_dcnv3_smallk_fused(const void *x, const void *offset, const void *mask_logits, const float *w2, const float *bias, void *out,
                          size_t offset_elems, size_t mask_elems, const gp_dcnv3_desc *d, int K, int C_out, int dtype, void *stream) {
    if (!x || !offset || !mask_logits || !w2 || !bias || !out || !d) return GP_ERR_NULL;
    if (K != 3 || C_out != 256 || d->G != 4 || d->kh != 3 || d->kw != 3 || d->remove_center) return GP_ERR_UNSUPPORTED;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->sh <= 0 || d->sw <= 0 || d->dh <= 0 || d->dw <= 0) return GP_ERR_SHAPE;
    if (d->Ho != gp_dcnv3_out_size(d->H, d->kh, d->sh, d->ph, d->dh) || d->Wo != gp_dcnv3_out_size(d->W, d->kw, d->sw, d->pw, d->dw)) return GP_ERR_SHAPE;
    const long long n_pix = (long long)d->N * d->Ho * d->Wo;
    if (offset_elems < (size_t)n_pix * 4 * 9 * 2 || mask_elems < (size_t)n_pix * 4 * 9) return GP_ERR_SHAPE;   // flat-prefix addressing
    if (!al16(out) || !al16(w2) || (reinterpret_cast<uintptr_t>(offset) & 7u)) return GP_ERR_ALIGN;
    if (n_pix == 0) return GP_OK;
    SmallKFusedParams p;
    p.H = d->H; p.W = d->W; p.Ho = d->Ho; p.Wo = d->Wo; p.sh = d->sh; p.sw = d->sw; p.dh = d->dh; p.dw = d->dw;
    p.half_h = (d->dh * (d->kh - 1)) >> 1; p.half_w = (d->dw * (d->kw - 1)) >> 1;
    p.base_h = p.half_h - d->ph; p.base_w = p.half_w - d->pw;
    p.scale = d->offset_scale; p.n_pix = n_pix;
    constexpr int PX = 4;
    const long long want = (n_pix + 8 * PX - 1) / (8 * PX);
    const unsigned grid = (unsigned)(want < 148ll * 3 ? want : 148ll * 3);   // 3 CTAs / SM (launch bounds), persistent over pixel groups
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case GP_F32: dcnv3_smallk_fused_kernel<float, PX><<<grid, 256, 0, st>>>((const float *)x, (const float *)offset, (const float *)mask_logits, w2, bias, (float *)out, p); break;
        case GP_BF16: dcnv3_smallk_fused_kernel<__nv_bfloat16, PX><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (const __nv_bfloat16 *)offset, (const __nv_bfloat16 *)mask_logits, w2, bias, (__nv_bfloat16 *)out, p); break;
        case GP_F16: dcnv3_smallk_fused_kernel<__half, PX><<<grid, 256, 0, st>>>((const __half *)x, (const __half *)offset, (const __half *)mask_logits, w2, bias, (__half *)out, p); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

size_t gp_groupnorm_backward_workspace_floats(int N, int H, int W, int C, int G) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || G <= 0) return 0;
    int ppc = 0;
    const dim3 sgrid = slab_grid(N, H * W, 8, &ppc);
    return (size_t)N * G * 2 + (size_t)N * sgrid.x * C * 2;   // (A, B)/cnt per (n, g) followed by the per-channel slab partials
}

int gp_groupnorm_act_backward(const void *x, const void *dy, const float *stats, const float *gamma, const float *beta, void *dx,
                              float *dgamma, float *dbeta, float *ws, size_t ws_floats, int N, int H, int W, int C, int G, int act,
                              int dtype, void *stream) {
    if (!x || !dy || !stats || !gamma || !beta || !dx || !dgamma || !dbeta || !ws) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || C <= 0 || G <= 0 || C % G || (C / G) % 4 || C / 4 > 256) return GP_ERR_SHAPE;
    if (act < 0 || act > 2) return GP_ERR_UNSUPPORTED;
    if (dtype != GP_F32 && C % 8) return GP_ERR_SHAPE;
    if (!al16(x) || !al16(dy) || !al16(dx)) return GP_ERR_ALIGN;
    if (ws_floats < gp_groupnorm_backward_workspace_floats(N, H, W, C, G)) return GP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    int sppc = 0, appc = 0;
    const dim3 sgrid = slab_grid(N, HW, 8, &sppc);
    const dim3 agrid = slab_grid(N, HW, 16, &appc);
    float *gstat = ws, *partial = ws + (size_t)N * G * 2;
#define GP_GNB(TT, AA) do { \
        gn_bwd_stats_kernel<TT, AA><<<sgrid, 256, 256 * 8 * sizeof(float), st>>>((const TT *)x, (const TT *)dy, stats, gamma, beta, partial, HW, C, G, sppc); \
        gn_bwd_group_kernel<<<(N * G + 127) / 128, 128, 0, st>>>(partial, gamma, gstat, N * G, G, C, (int)sgrid.x, 1.f / ((float)HW * (C / G))); \
        gn_bwd_param_kernel<<<(C + 7) / 8, 256, 0, st>>>(partial, dgamma, dbeta, C, N * (int)sgrid.x); \
        gn_bwd_apply_kernel<TT, AA><<<agrid, 256, 0, st>>>((const TT *)x, (const TT *)dy, stats, gstat, gamma, beta, (TT *)dx, HW, C, G, appc); } while (0)
#define GP_GNB_T(TT) do { if (act == ACT_RELU) GP_GNB(TT, ACT_RELU); else if (act == ACT_GELU) GP_GNB(TT, ACT_GELU); else GP_GNB(TT, ACT_NONE); } while (0)
    switch (dtype) {
        case GP_F32: GP_GNB_T(float); break;
        case GP_BF16: GP_GNB_T(__nv_bfloat16); break;
        case GP_F16: GP_GNB_T(__half); break;
        default: return GP_ERR_DTYPE;
    }
#undef GP_GNB_T
#undef GP_GNB
    count_launch(4);
    return (int)cudaGetLastError();
}

static int groupnorm_act_conv1x1_impl(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta,
                                      const float *w, const float *bias, int N, int H, int W, int C, int G, float eps, int act, int OC,
                                      int dtype, void *stream, bool have_stats) {
    if (!x || !y || !stats || !gamma || !beta || !w || !bias) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || G <= 0 || C % G || (C / G) % 4) return GP_ERR_SHAPE;
    if (C != 256 || OC != 3 || act < 0 || act > 2) return GP_ERR_UNSUPPORTED;   // the decoder's out_layer
    if (!al16(x) || !al16(gamma) || !al16(beta)) return GP_ERR_ALIGN;
    if (stats_floats < (have_stats ? (size_t)N * G * 2 : gp_groupnorm_workspace_floats(N, H, W, G))) return GP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H * W;
    int appc = 0;
    const dim3 agrid = slab_grid(N, HW, 8, &appc);
#define GP_GNC_APPLY(TT, AA) gn_act_conv1x1_kernel<TT, AA, 8, 3><<<agrid, 256, 0, st>>>((const TT *)x, stats, gamma, beta, w, bias, (TT *)y, HW, G, eps, appc)
#define GP_GNC(TT) do { if (!have_stats) gn_statistics<TT>(x, stats, N, HW, C, G, eps, st); \
                        if (act == ACT_RELU) GP_GNC_APPLY(TT, ACT_RELU); else if (act == ACT_GELU) GP_GNC_APPLY(TT, ACT_GELU); \
                        else GP_GNC_APPLY(TT, ACT_NONE); } while (0)
    switch (dtype) {
        case GP_F32: GP_GNC(float); break;
        case GP_BF16: GP_GNC(__nv_bfloat16); break;
        case GP_F16: GP_GNC(__half); break;
        default: return GP_ERR_DTYPE;
    }
#undef GP_GNC
#undef GP_GNC_APPLY
    count_launch();
    return (int)cudaGetLastError();
}

int gp_groupnorm_act_conv1x1(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta,
                             const float *w, const float *bias, int N, int H, int W, int C, int G, float eps, int act, int OC,
                             int dtype, void *stream) {
    return groupnorm_act_conv1x1_impl(x, y, stats, stats_floats, gamma, beta, w, bias, N, H, W, C, G, eps, act, OC, dtype, stream, false);
}

int gp_groupnorm_apply_conv1x1(const void *x, void *y, const float *stats, size_t stats_floats, const float *gamma, const float *beta,
                               const float *w, const float *bias, int N, int H, int W, int C, int G, float eps, int act, int OC,
                               int dtype, void *stream) {
    return groupnorm_act_conv1x1_impl(x, y, const_cast<float *>(stats), stats_floats, gamma, beta, w, bias, N, H, W, C, G, eps, act, OC, dtype,
                                      stream, true);
}

int gp_upsample_bilinear2x(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream) {
    if (!x || !y) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || C <= 0 || C % 4 || C / 4 > 256) return GP_ERR_SHAPE;
    if (dtype != GP_F32 && C % 8) return GP_ERR_SHAPE;   // 16-byte channel vectors
    if (!al16(x) || !al16(y)) return GP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    {   // column-strip kernel when the shape tiles exactly (the decoder: C = 256, Wo = 32 / 64)
        const int V = dtype == GP_F32 ? 4 : 8, q = C / V;
        if (q <= 256 && 256 % q == 0 && (2 * W) % (256 / q) == 0 && N <= 65535) {
            const int cols = 256 / q, gx = 2 * W / cols, Ho = 2 * H;
            int slabs = (int)((148ll * 16 + (long long)N * gx - 1) / ((long long)N * gx));   // >= 16 CTAs per SM worth of work
            if (slabs > Ho / 4) slabs = Ho / 4 > 0 ? Ho / 4 : 1;                              // >= 4 rows per CTA
            if (slabs < 1) slabs = 1;
            const int rpc = (Ho + slabs - 1) / slabs;
            const dim3 grid((unsigned)gx, (unsigned)((Ho + rpc - 1) / rpc), (unsigned)N);
            switch (dtype) {
                case GP_F32: upsample2x_strip_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, H, W, C, rpc); break;
                case GP_BF16: upsample2x_strip_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, H, W, C, rpc); break;
                case GP_F16: upsample2x_strip_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, (__half *)y, H, W, C, rpc); break;
                default: return GP_ERR_DTYPE;
            }
            count_launch();
            return (int)cudaGetLastError();
        }
    }
    int ppc = 0;
    const dim3 grid = slab_grid(N, 4 * H * W, 16, &ppc);
    switch (dtype) {
        case GP_F32: upsample2x_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, H, W, C, ppc); break;
        case GP_BF16: upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, H, W, C, ppc); break;
        case GP_F16: upsample2x_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, (__half *)y, H, W, C, ppc); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_upsample_bilinear2x_backward(const void *dy, void *dx, int N, int H, int W, int C, int dtype, void *stream) {
    if (!dy || !dx) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || C <= 0 || C % 4 || C / 4 > 256) return GP_ERR_SHAPE;
    if (dtype != GP_F32 && C % 8) return GP_ERR_SHAPE;
    if (!al16(dy) || !al16(dx)) return GP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int ppc = 0;
    const dim3 grid = slab_grid(N, H * W, 16, &ppc);
    switch (dtype) {
        case GP_F32: upsample2x_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)dy, (float *)dx, H, W, C, ppc); break;
        case GP_BF16: upsample2x_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)dy, (__nv_bfloat16 *)dx, H, W, C, ppc); break;
        case GP_F16: upsample2x_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half *)dy, (__half *)dx, H, W, C, ppc); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_bias_add_relu(void *y, const void *residual, const float *bias, long long rows, int C, int dtype, void *stream) {
    if (!y || !residual || !bias) return GP_ERR_NULL;
    if (rows < 0 || C <= 0 || C % 8) return GP_ERR_SHAPE;
    if (!al16(y) || !al16(residual)) return GP_ERR_ALIGN;
    if (rows == 0) return GP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int V = dtype == GP_F32 ? 4 : 8;
    const long long n_vec = rows * C / V, want = (n_vec + 255) / 256;
    const unsigned grid = (unsigned)(want < 148ll * 32 ? want : 148ll * 32);
    switch (dtype) {
        case GP_F32: bias_add_relu_kernel<float><<<grid, 256, 0, st>>>((float *)y, (const float *)residual, bias, n_vec, C); break;
        case GP_BF16: bias_add_relu_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16 *)y, (const __nv_bfloat16 *)residual, bias, n_vec, C); break;
        case GP_F16: bias_add_relu_kernel<__half><<<grid, 256, 0, st>>>((__half *)y, (const __half *)residual, bias, n_vec, C); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_stem_s2d_pack(const float *img, void *out, int N, int H, int W, int dtype, void *stream) {
    if (!img || !out) return GP_ERR_NULL;
    if (N <= 0 || H <= 0 || W <= 0 || H % 2 || W % 2) return GP_ERR_SHAPE;
    if (!al16(out) || (reinterpret_cast<uintptr_t>(img) & 7u)) return GP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * (H / 2 + 3) * (W / 2 + 3);
    const long long want = (total + 255) / 256;
    const unsigned grid = (unsigned)(want < 148ll * 32 ? want : 148ll * 32);
    switch (dtype) {
        case GP_F32: stem_s2d_pack_kernel<float><<<grid, 256, 0, st>>>(img, (float *)out, N, H, W); break;
        case GP_BF16: stem_s2d_pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(img, (__nv_bfloat16 *)out, N, H, W); break;
        case GP_F16: stem_s2d_pack_kernel<__half><<<grid, 256, 0, st>>>(img, (__half *)out, N, H, W); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_maxpool3x3s2(const void *x, void *y, int N, int H, int W, int C, int relu, int dtype, void *stream) {
    if (!x || !y) return GP_ERR_NULL;
    if (N <= 0 || N > 65535 || H <= 0 || W <= 0 || C <= 0 || C % 4 || C / 4 > 256) return GP_ERR_SHAPE;
    if (dtype != GP_F32 && C % 8) return GP_ERR_SHAPE;   // 16-byte channel vectors
    if (!al16(x) || !al16(y)) return GP_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    int ppc = 0;
    const dim3 grid = slab_grid(N, Ho * Wo, 16, &ppc);
    switch (dtype) {
        case GP_F32: maxpool3x3s2_kernel<float><<<grid, 256, 0, st>>>((const float *)x, (float *)y, H, W, C, ppc, relu ? 0.f : -INFINITY); break;
        case GP_BF16: maxpool3x3s2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, (__nv_bfloat16 *)y, H, W, C, ppc, relu ? 0.f : -INFINITY); break;
        case GP_F16: maxpool3x3s2_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, (__half *)y, H, W, C, ppc, relu ? 0.f : -INFINITY); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_mhsa_tokens(const void *qkv, void *out, int B, int NT, int NH, int HD, float scale, int dtype, void *stream) {
    if (!qkv || !out) return GP_ERR_NULL;
    if (B < 0 || B > 65535 || NH <= 0 || NH > 65535) return GP_ERR_SHAPE;
    if (NT != 64 || HD != 32) return GP_ERR_UNSUPPORTED;   // MAPTransformerEncoer: 8x8 patches, 256 / 8 heads
    if (B == 0) return GP_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(NH, B);
    switch (dtype) {
        case GP_F32: mhsa_tokens_kernel<float, 64, 32><<<grid, 64, 0, st>>>((const float *)qkv, (float *)out, NH, scale); break;
        case GP_BF16: mhsa_tokens_kernel<__nv_bfloat16, 64, 32><<<grid, 64, 0, st>>>((const __nv_bfloat16 *)qkv, (__nv_bfloat16 *)out, NH, scale); break;
        case GP_F16: mhsa_tokens_kernel<__half, 64, 32><<<grid, 64, 0, st>>>((const __half *)qkv, (__half *)out, NH, scale); break;
        default: return GP_ERR_DTYPE;
    }
    count_launch();
    return (int)cudaGetLastError();
}

int gp_pose_decode(const float *rot6, const float *t, const float *cam, int cam_batched, const float *centers, const float *whs,
                   const float *ratios, float *rot_out, float *trans_out, int B, int is_allo, float z_calib, void *stream) {
    if (!rot6 || !t || !cam || !centers || !whs || !ratios || !rot_out || !trans_out) return GP_ERR_NULL;
    if (B < 0) return GP_ERR_SHAPE;
    if (B == 0) return GP_OK;
    pose_decode_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rot6, t, cam, cam_batched ? 9 : 0, centers, whs, ratios,
                                                                         rot_out, trans_out, B, is_allo, z_calib);
    count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
