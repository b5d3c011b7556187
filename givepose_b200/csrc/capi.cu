// givepose_b200 -- C-ABI entry points (include/givepose_b200.h): argument checks, kernel selection,
// launches.  Host-side counterpart of the reference's network/ops_dcnv3/src/cuda/dcnv3_cuda.cu:21-174,
// without the im2col_step chunk loop (the flat offset/mask addressing makes chunking a no-op) and without
// the at::zeros memset of an output that is fully overwritten (dcnv3_cuda.cu:55-57).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "dcnv3_kernels.cuh"
#include "dcnv3_gin_binned.cuh"
#include "dcnv3_bwd_fused.cuh"
#include "dcnv3_fwd_rows.cuh"
#include "tc_common.cuh"

namespace gp {
// The file is compiled twice (Makefile): GP_PART 1 = forward kernels + option / host-buffer entry points, GP_PART 2 = backward
// kernels; 0 (default) = everything in one object.  Template instantiation of the sampling kernels dominates the build time.
#ifndef GP_PART
#define GP_PART 0
#endif
#if GP_PART != 2
unsigned long long g_launches = 0;
#endif

struct Tuning {
    int tile_h = 8, tile_w = 8, gs = 2;   // measured best on B200 (profiles/r01_sweep.md)
    int vec16 = 8;   // forward: channels per lane for 16-bit storage (8 = 16-byte requests, 4 = 8-byte requests)
                     // backward always uses 4: its fp32 reductions then cover whole 128-byte lines per request
    int bwd_mode = 0;              // 0: one-pass scatter kernel (fastest measured, profiles/r02_it1_*); 1: grad_offset/grad_mask kernel + binned grad_input kernel
    int gin_th = 8, gin_tw = 8;    // output tile of the binned grad_input kernel
    int gin_nt = 192;              // its CTA size
    int fwd_mode = 1;
    int last_fwd_kernel = -1;      // read-only (GP_OPT_LAST_FWD_KERNEL): 0 dcnv3_fwd_tile, 1 dcnv3_fwd_rows, 2 dcnv3_fwd_generic
    bool init = false;
};
#if GP_PART != 2
Tuning g_tune;
#else
extern Tuning g_tune;
#endif

static void init_tuning() {
    if (g_tune.init) return;
    g_tune.init = true;
    if (const char *e = getenv("GP_TILE_H")) g_tune.tile_h = atoi(e);
    if (const char *e = getenv("GP_TILE_W")) g_tune.tile_w = atoi(e);
    if (const char *e = getenv("GP_GS")) g_tune.gs = atoi(e);
    if (const char *e = getenv("GP_VEC16")) g_tune.vec16 = atoi(e) == 4 ? 4 : 8;
    if (const char *e = getenv("GP_BWD_MODE")) { const int v = atoi(e); g_tune.bwd_mode = v >= 0 && v <= 2 ? v : 0; }
    if (const char *e = getenv("GP_FWD_MODE")) g_tune.fwd_mode = atoi(e) ? 1 : 0;
}

static int make_params(const gp_dcnv3_desc *d, KParams &p) {
    if (!d) return GP_ERR_NULL;
    if (d->N <= 0 || d->H <= 0 || d->W <= 0 || d->G <= 0 || d->gc <= 0 || d->kh <= 0 || d->kw <= 0 || d->sh <= 0 ||
        d->sw <= 0 || d->dh <= 0 || d->dw <= 0 || d->ph < 0 || d->pw < 0)
        return GP_ERR_SHAPE;
    if (d->remove_center != 0 && d->remove_center != 1) return GP_ERR_SHAPE;
    if (d->remove_center && (d->kh % 2 == 0 || d->kw % 2 == 0 || d->kh != d->kw)) return GP_ERR_UNSUPPORTED;
    if (d->Ho != gp_dcnv3_out_size(d->H, d->kh, d->sh, d->ph, d->dh) ||
        d->Wo != gp_dcnv3_out_size(d->W, d->kw, d->sw, d->pw, d->dw) || d->Ho <= 0 || d->Wo <= 0)
        return GP_ERR_SHAPE;
    memset(&p, 0, sizeof(p));
    p.N = d->N; p.H = d->H; p.W = d->W; p.G = d->G; p.gc = d->gc; p.C = d->G * d->gc;
    p.kh = d->kh; p.kw = d->kw; p.sh = d->sh; p.sw = d->sw; p.ph = d->ph; p.pw = d->pw; p.dh = d->dh; p.dw = d->dw;
    p.remove_center = d->remove_center;
    p.P = d->kh * d->kw - d->remove_center;
    p.Ho = d->Ho; p.Wo = d->Wo;
    p.half_h = (d->dh * (d->kh - 1)) >> 1;
    p.half_w = (d->dw * (d->kw - 1)) >> 1;
    p.base_h = p.half_h - d->ph;
    p.base_w = p.half_w - d->pw;
    p.scale = d->offset_scale;
    p.n_units = (long long)d->N * d->Ho * d->Wo * d->G;
    if (p.P <= 0) return GP_ERR_SHAPE;
    p.off_q = (long long)p.G * p.P * 2;
    p.msk_q = (long long)p.G * p.P;
    return GP_OK;
}

static size_t elem_size(int dtype) {
    switch (dtype) {
        case GP_F32: return 4;
        case GP_BF16: case GP_F16: return 2;
        case GP_F64: return 8;
    }
    return 0;
}

static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// can the tiled/vectorised kernels take this call?  fills the tiling fields of p.
static bool plan_tiled(KParams &p, int dtype, int *L_out, int *vec_out, bool backward = false) {
    if (dtype == GP_F64) return false;
    init_tuning();
    int vec = dtype == GP_F32 ? 4 : (backward && p.gc % 4 == 0 && p.gc / 4 <= 32 ? 4 : g_tune.vec16);
    if (vec == 8 && (p.gc % 8 || p.gc / 8 > 32) && p.gc % 4 == 0) vec = 4;
    if (p.gc % vec) return false;
    const int L = p.gc / vec;
    if (L > 32 || (L & (L - 1))) return false;
    if ((long long)p.H * p.W * p.C * 4 >= (1ll << 31)) return false;   // 32-bit byte offsets inside one image
    int th = g_tune.tile_h, tw = g_tune.tile_w, gs = g_tune.gs;
    auto pow2_floor = [](int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; };
    auto lg2 = [](int v) { int r = 0; while ((1 << r) < v) ++r; return r; };
    th = pow2_floor(th < 1 ? 1 : th);
    tw = pow2_floor(tw < 1 ? 1 : tw);
    gs = pow2_floor(gs < 1 ? 1 : gs);
    while (gs > 1 && p.G % gs) gs /= 2;
    // keep the sampling records under 48 KB of shared memory
    while ((long long)th * tw * gs * (p.P * 24 + 12) > 48 * 1024) {
        if (gs > 1) { gs /= 2; continue; }
        if (th >= tw && th > 1) th /= 2; else if (tw > 1) tw /= 2; else return false;
    }
    p.tile_h = th; p.tile_w = tw; p.gs = gs; p.gchunks = p.G / gs;
    p.lg_tw = lg2(tw); p.lg_tp = lg2(th * tw); p.lg_gs = lg2(gs);
    p.tiles_y = (p.Ho + th - 1) / th;
    p.tiles_x = (p.Wo + tw - 1) / tw;
    const long long ctas = (long long)p.N * p.tiles_y * p.tiles_x * p.gchunks;
    if (ctas >= (1ll << 31)) return false;
    *L_out = L;
    *vec_out = vec;
    return true;
}

static size_t tile_smem(const KParams &p, bool softmax) {
    (void)softmax;   // the fused softmax lives in the builder thread's registers
    return (size_t)p.tile_h * p.tile_w * p.gs * (p.P * 24 + 4);
}
static unsigned tile_grid(const KParams &p) { return (unsigned)((long long)p.N * p.tiles_y * p.tiles_x * p.gchunks); }

#if GP_PART != 2
template <typename T, int VEC, bool SOFTMAX>
static void launch_fwd_tile(const void *in, const void *off, const void *msk, void *out, const KParams &p, int L,
                            cudaStream_t st) {
    const bool k3 = p.P == 9;
    const size_t sm = tile_smem(p, SOFTMAX);
    const unsigned grid = tile_grid(p);
#define GP_FWD(LL, K3)                                                                                         \
    dcnv3_fwd_tile<T, VEC, LL, K3, SOFTMAX><<<grid, kTileThreads, sm, st>>>((const T *)in, (const T *)off,     \
                                                                            (const T *)msk, (T *)out, p)
#define GP_FWD_L(LL) do { if (k3) GP_FWD(LL, true); else GP_FWD(LL, false); } while (0)
    switch (L) {
        case 1: GP_FWD_L(1); break;
        case 2: GP_FWD_L(2); break;
        case 4: GP_FWD_L(4); break;
        case 8: GP_FWD_L(8); break;
        case 16: GP_FWD_L(16); break;
        case 32: GP_FWD_L(32); break;
    }
#undef GP_FWD_L
#undef GP_FWD
    count_launch();
}

#endif
#if GP_PART != 1
template <typename T, int VEC, bool GIN>
static void launch_bwd_tile(const void *in, const void *off, const void *msk, const void *gout, float *gin, void *goff,
                            void *gmsk, const KParams &p, int L, cudaStream_t st) {
    const bool k3 = p.P == 9;
    const size_t sm = tile_smem(p, false);
    const unsigned grid = tile_grid(p);
#define GP_BWD(LL, K3)                                                                                          \
    dcnv3_bwd_tile<T, VEC, LL, K3, GIN><<<grid, kTileThreads, sm, st>>>((const T *)in, (const T *)off, (const T *)msk, \
                                                                   (const T *)gout, gin, (T *)goff, (T *)gmsk, p)
#define GP_BWD_L(LL) do { if (k3) GP_BWD(LL, true); else GP_BWD(LL, false); } while (0)
    switch (L) {
        case 1: GP_BWD_L(1); break;
        case 2: GP_BWD_L(2); break;
        case 4: GP_BWD_L(4); break;
        case 8: GP_BWD_L(8); break;
        case 16: GP_BWD_L(16); break;
        case 32: GP_BWD_L(32); break;
    }
#undef GP_BWD_L
#undef GP_BWD
    count_launch();
}

// ---- binned grad_input (dcnv3_gin_binned.cuh) ----------------------------------------------------------------------
// fills the tiling fields of pg (one group per CTA) and the CTA size; false: this call takes the scatter path
static bool plan_gin_binned(const KParams &p, int dtype, KParams &pg, int *L_out, int *nt_out) {
    if (dtype == GP_F64 || p.gc % 4) return false;
    const int L = p.gc / 4;
    if (L > 32 || (L & (L - 1))) return false;
    if (p.H > 32767 || p.W > 32767) return false;
    if ((long long)p.H * p.W * p.C * 4 >= (1ll << 31)) return false;
    init_tuning();
    int nt = g_tune.gin_nt;
    if (nt != 128 && nt != 256) nt = 192;
    if ((L != 8 && L != 16) || nt % L) nt = 192;        // the other CTA sizes are only instantiated for L = 8, 16
    const int spt = nt == 128 ? 5 : 3;
    auto pow2_floor = [](int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; };
    auto lg2 = [](int v) { int r = 0; while ((1 << r) < v) ++r; return r; };
    int th = pow2_floor(g_tune.gin_th < 1 ? 1 : g_tune.gin_th), tw = pow2_floor(g_tune.gin_tw < 1 ? 1 : g_tune.gin_tw);
    while (th * tw * p.P > nt * spt) {
        if (th >= tw && th > 1) th /= 2; else if (tw > 1) tw /= 2; else return false;
    }
    pg = p;
    pg.tile_h = th; pg.tile_w = tw; pg.gs = 1; pg.gchunks = p.G;
    pg.lg_tw = lg2(tw); pg.lg_tp = lg2(th * tw); pg.lg_gs = 0;
    pg.tiles_y = (p.Ho + th - 1) / th;
    pg.tiles_x = (p.Wo + tw - 1) / tw;
    if ((long long)p.N * pg.tiles_y * pg.tiles_x * pg.gchunks >= (1ll << 31)) return false;
    *L_out = L;
    *nt_out = nt;
    return true;
}

template <typename T, int L, int NT, int SPT, int MINB>
static cudaError_t launch_gin_one(const void *off, const void *msk, const void *gout, float *gin, const KParams &pg,
                                  cudaStream_t st) {
    const size_t sm = gin_binned_smem<NT, SPT>(pg.tile_h * pg.tile_w, L);
    const unsigned grid = tile_grid(pg);
    cudaError_t e = cudaSuccess;
    if (pg.P == 9) {
        auto k = dcnv3_gin_binned<T, L, NT, SPT, MINB, true>;
        if (sm > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k<<<grid, NT, sm, st>>>((const T *)off, (const T *)msk, (const T *)gout, gin, pg);
    } else {
        auto k = dcnv3_gin_binned<T, L, NT, SPT, MINB, false>;
        if (sm > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k<<<grid, NT, sm, st>>>((const T *)off, (const T *)msk, (const T *)gout, gin, pg);
    }
    count_launch();
    return e;
}

template <typename T>
static cudaError_t launch_gin_binned(const void *off, const void *msk, const void *gout, float *gin, const KParams &pg,
                                     int L, int nt, cudaStream_t st) {
    if (nt == 128 && L == 8) return launch_gin_one<T, 8, 128, 5, 8>(off, msk, gout, gin, pg, st);
    if (nt == 128 && L == 16) return launch_gin_one<T, 16, 128, 5, 8>(off, msk, gout, gin, pg, st);
    if (nt == 256 && L == 8) return launch_gin_one<T, 8, 256, 3, 4>(off, msk, gout, gin, pg, st);
    if (nt == 256 && L == 16) return launch_gin_one<T, 16, 256, 3, 4>(off, msk, gout, gin, pg, st);
    switch (L) {
        case 1: return launch_gin_one<T, 1, 192, 3, 5>(off, msk, gout, gin, pg, st);
        case 2: return launch_gin_one<T, 2, 192, 3, 5>(off, msk, gout, gin, pg, st);
        case 4: return launch_gin_one<T, 4, 192, 3, 5>(off, msk, gout, gin, pg, st);
        case 8: return launch_gin_one<T, 8, 192, 3, 5>(off, msk, gout, gin, pg, st);
        case 16: return launch_gin_one<T, 16, 192, 3, 5>(off, msk, gout, gin, pg, st);
        case 32: return launch_gin_one<T, 32, 192, 3, 5>(off, msk, gout, gin, pg, st);
    }
    return cudaErrorInvalidValue;
}

// ---- fused backward with in-SM pre-aggregation (dcnv3_bwd_fused.cuh) --------------------------------------------------
static bool plan_bwd_fused(const KParams &p, int dtype, KParams &pf, int *L_out) {
    if (dtype == GP_F64 || p.gc % 4) return false;
    const int L = p.gc / 4;
    if (L > 32 || (L & (L - 1))) return false;
    if (p.H > 32767 || p.W > 32767) return false;
    if ((long long)p.H * p.W * p.C * 4 >= (1ll << 31)) return false;
    init_tuning();
    auto pow2_floor = [](int v) { int r = 1; while (r * 2 <= v) r *= 2; return r; };
    auto lg2 = [](int v) { int r = 0; while ((1 << r) < v) ++r; return r; };
    int th = pow2_floor(g_tune.gin_th < 1 ? 1 : g_tune.gin_th), tw = pow2_floor(g_tune.gin_tw < 1 ? 1 : g_tune.gin_tw);
    while (th * tw * p.P > kFusedThreads * kFusedSPT || th * tw > 256) {
        if (th >= tw && th > 1) th /= 2; else if (tw > 1) tw /= 2; else return false;
    }
    pf = p;
    pf.tile_h = th; pf.tile_w = tw; pf.gs = 1; pf.gchunks = p.G;
    pf.lg_tw = lg2(tw); pf.lg_tp = lg2(th * tw); pf.lg_gs = 0;
    pf.tiles_y = (p.Ho + th - 1) / th;
    pf.tiles_x = (p.Wo + tw - 1) / tw;
    if ((long long)p.N * pf.tiles_y * pf.tiles_x * pf.gchunks >= (1ll << 31)) return false;
    if (bwd_fused_smem(th * tw, p.P, L) > 160 * 1024) return false;
    *L_out = L;
    return true;
}

template <typename T, int L>
static cudaError_t launch_bwd_fused_one(const void *in, const void *off, const void *msk, const void *gout, float *gin,
                                        void *goff, void *gmsk, const KParams &pf, cudaStream_t st) {
    const size_t sm = bwd_fused_smem(pf.tile_h * pf.tile_w, pf.P, L);
    const unsigned grid = tile_grid(pf);
    cudaError_t e = cudaSuccess;
    if (pf.P == 9) {
        auto k = dcnv3_bwd_fused<T, L, true, 6>;
        if (sm > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k<<<grid, kFusedThreads, sm, st>>>((const T *)in, (const T *)off, (const T *)msk, (const T *)gout, gin, (T *)goff, (T *)gmsk, pf);
    } else {
        auto k = dcnv3_bwd_fused<T, L, false, 6>;
        if (sm > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e == cudaSuccess) k<<<grid, kFusedThreads, sm, st>>>((const T *)in, (const T *)off, (const T *)msk, (const T *)gout, gin, (T *)goff, (T *)gmsk, pf);
    }
    count_launch();
    return e;
}

template <typename T>
static cudaError_t launch_bwd_fused(const void *in, const void *off, const void *msk, const void *gout, float *gin, void *goff,
                                    void *gmsk, const KParams &pf, int L, cudaStream_t st) {
    switch (L) {
        case 1: return launch_bwd_fused_one<T, 1>(in, off, msk, gout, gin, goff, gmsk, pf, st);
        case 2: return launch_bwd_fused_one<T, 2>(in, off, msk, gout, gin, goff, gmsk, pf, st);
        case 4: return launch_bwd_fused_one<T, 4>(in, off, msk, gout, gin, goff, gmsk, pf, st);
        case 8: return launch_bwd_fused_one<T, 8>(in, off, msk, gout, gin, goff, gmsk, pf, st);
        case 16: return launch_bwd_fused_one<T, 16>(in, off, msk, gout, gin, goff, gmsk, pf, st);
        case 32: return launch_bwd_fused_one<T, 32>(in, off, msk, gout, gin, goff, gmsk, pf, st);
    }
    return cudaErrorInvalidValue;
}

#endif
#if GP_PART != 2
template <typename T, bool SOFTMAX>
static void launch_fwd_generic(const void *in, const void *off, const void *msk, void *out, const KParams &p,
                               cudaStream_t st) {
    const long long total = p.n_units * p.gc;
    const long long blocks = (total + 255) / 256;
    const unsigned grid = (unsigned)(blocks > (1ll << 20) ? (1ll << 20) : blocks);
    dcnv3_fwd_generic<T, SOFTMAX><<<grid, 256, 0, st>>>((const T *)in, (const T *)off, (const T *)msk, (T *)out, p);
    count_launch();
}

#endif
#if GP_PART != 1
template <typename T>
static void launch_bwd_generic(const void *in, const void *off, const void *msk, const void *gout, void *gin_acc,
                               void *goff, void *gmsk, const KParams &p, cudaStream_t st) {
    const unsigned grid = (unsigned)(p.n_units > (1ll << 20) ? (1ll << 20) : p.n_units);
    dcnv3_bwd_generic<T><<<grid, 64, 0, st>>>((const T *)in, (const T *)off, (const T *)msk, (const T *)gout,
                                              (typename AccOf<T>::type *)gin_acc, (T *)goff, (T *)gmsk, p);
    count_launch();
}

#endif
#if GP_PART != 2
// ---- forward with TMA-staged offset / mask rows (dcnv3_fwd_rows.cuh): 3x3 kernels ------------------------------------------
// 3-D row tensor {row values, Wo, N*Ho} of `pitch` elements per pixel; box = {bw values, tile_w, tile_h}
static bool make_rows_map(CUtensorMap *map, const void *base, int dtype, long long inner, long long pitch, const KParams &p, int bw) {
    tc::EncodeTiledFn fn = tc::encode_fn();
    if (!fn) return false;
    const size_t es = elem_size(dtype);
    const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)p.Wo, (cuuint64_t)p.N * p.Ho};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * es, (cuuint64_t)pitch * es * p.Wo};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)p.tile_w, (cuuint32_t)p.tile_h};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dt = dtype == GP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : dtype == GP_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    return fn(map, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// true: launched.  false: this call keeps the per-thread row reads of dcnv3_fwd_tile (shape / alignment outside the TMA rules).
template <typename T, int VEC, bool SOFTMAX>
static bool launch_fwd_rows(const void *in, const void *off, const void *msk, void *out, const KParams &p, int L, int dtype,
                            cudaStream_t st) {
    if (p.P != 9 || p.remove_center || (L != 4 && L != 8 && L != 16)) return false;
    const size_t es = sizeof(T);
    if ((p.off_q * es) % 16 || (p.msk_q * es) % 16 || !aligned16(off) || !aligned16(msk)) return false;   // TMA: 16-byte strides / base
    if ((long long)p.N * p.Ho >= (1ll << 31) || p.tile_w > 256 || p.tile_h > 256) return false;
    const int per16 = (int)(16 / es);
    // box widths: the CTA's gs*18 / gs*9 values + up to per16-1 leading values (boxes start on 16-byte boundaries), in 16-byte units
    const int bw_off = (p.gs * 18 + 2 * (per16 - 1)) / per16 * per16, bw_msk = (p.gs * 9 + 2 * (per16 - 1)) / per16 * per16;
    if (bw_off > 256) return false;
    const size_t sm = fwd_rows_smem(p.tile_h * p.tile_w, p.gs, bw_off, bw_msk, (int)es);
    if (sm > 48 * 1024) return false;
    CUtensorMap mo, mm;
    if (!make_rows_map(&mo, off, dtype, (long long)p.G * 18, p.off_q, p, bw_off) ||
        !make_rows_map(&mm, msk, dtype, (long long)p.G * 9, p.msk_q, p, bw_msk))
        return false;
    const unsigned grid = tile_grid(p);
#define GP_FWDR(LL) dcnv3_fwd_rows<T, VEC, LL, SOFTMAX><<<grid, kTileThreads, sm, st>>>((const T *)in, (T *)out, mo, mm, p, bw_off, bw_msk)
    switch (L) {
        case 4: GP_FWDR(4); break;
        case 8: GP_FWDR(8); break;
        case 16: GP_FWDR(16); break;
    }
#undef GP_FWDR
    count_launch();
    return true;
}

// pitch > 0: `off` points at packed rows [G*P*2 offsets | G*P mask values | padding] of `pitch` elements per pixel
template <bool SOFTMAX>
static int forward_impl(const void *in, const void *off, const void *msk, void *out, const gp_dcnv3_desc *d, int dtype,
                        void *stream, long long pitch = 0) {
    if (!in || !off || !msk || !out) return GP_ERR_NULL;
    if (!elem_size(dtype)) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(d, p)) return e;
    if (!aligned16(in) || !aligned16(off) || !aligned16(out) || (!pitch && !aligned16(msk))) return GP_ERR_ALIGN;
    if (pitch) {
        if (pitch < (long long)p.G * p.P * 3 || pitch % 2) return GP_ERR_SHAPE;   // (w, h) pairs are read as one 2-element word
        p.off_q = p.msk_q = pitch;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int L = 0, vec = 0;
    if (pitch && !plan_tiled(p, dtype, &L, &vec)) return GP_ERR_UNSUPPORTED;   // packed rows: tiled kernels only
    if (plan_tiled(p, dtype, &L, &vec)) {
        bool done = false;
        if (g_tune.fwd_mode == 1) {
            if (dtype == GP_F32) done = launch_fwd_rows<float, 4, SOFTMAX>(in, off, msk, out, p, L, dtype, st);
            else if (dtype == GP_BF16 && vec == 8) done = launch_fwd_rows<__nv_bfloat16, 8, SOFTMAX>(in, off, msk, out, p, L, dtype, st);
            else if (dtype == GP_F16 && vec == 8) done = launch_fwd_rows<__half, 8, SOFTMAX>(in, off, msk, out, p, L, dtype, st);
        }
        g_tune.last_fwd_kernel = done ? 1 : 0;
        if (done) return (int)cudaGetLastError();
        if (dtype == GP_F32) launch_fwd_tile<float, 4, SOFTMAX>(in, off, msk, out, p, L, st);
        else if (dtype == GP_BF16 && vec == 8) launch_fwd_tile<__nv_bfloat16, 8, SOFTMAX>(in, off, msk, out, p, L, st);
        else if (dtype == GP_BF16) launch_fwd_tile<__nv_bfloat16, 4, SOFTMAX>(in, off, msk, out, p, L, st);
        else if (vec == 8) launch_fwd_tile<__half, 8, SOFTMAX>(in, off, msk, out, p, L, st);
        else launch_fwd_tile<__half, 4, SOFTMAX>(in, off, msk, out, p, L, st);
    } else {
        g_tune.last_fwd_kernel = 2;
        switch (dtype) {
            case GP_F32: launch_fwd_generic<float, SOFTMAX>(in, off, msk, out, p, st); break;
            case GP_BF16: launch_fwd_generic<__nv_bfloat16, SOFTMAX>(in, off, msk, out, p, st); break;
            case GP_F16: launch_fwd_generic<__half, SOFTMAX>(in, off, msk, out, p, st); break;
            case GP_F64: launch_fwd_generic<double, SOFTMAX>(in, off, msk, out, p, st); break;
        }
    }
    return (int)cudaGetLastError();
}

#endif
}  // namespace gp

using namespace gp;

extern "C" {

#if GP_PART != 2
int gp_abi_version(void) { return GP_ABI_VERSION; }

const char *gp_error_string(int code) {
    switch (code) {
        case GP_OK: return "ok";
        case GP_ERR_NULL: return "givepose_b200: required pointer is NULL";
        case GP_ERR_SHAPE: return "givepose_b200: inconsistent geometry (C != group*group_channels, bad dims or Ho/Wo)";
        case GP_ERR_DTYPE: return "givepose_b200: unknown dtype";
        case GP_ERR_ALIGN: return "givepose_b200: pointers must be 16-byte aligned";
        case GP_ERR_WORKSPACE: return "givepose_b200: workspace too small";
        case GP_ERR_UNSUPPORTED: return "givepose_b200: remove_center is only compatible with square odd kernel size";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "givepose_b200: unknown error";
}

int gp_dcnv3_out_size(int size, int k, int stride, int pad, int dil) {
    return (size + 2 * pad - (dil * (k - 1) + 1)) / stride + 1;
}

uint64_t gp_launch_count(void) { return g_launches; }
void gp_launch_count_reset(void) { g_launches = 0; }

int gp_set_tuning(int tile_h, int tile_w, int gs, int vec16) {
    init_tuning();
    if (tile_h > 0) g_tune.tile_h = tile_h;
    if (tile_w > 0) g_tune.tile_w = tile_w;
    if (gs > 0) g_tune.gs = gs;
    if (vec16 == 4 || vec16 == 8) g_tune.vec16 = vec16;
    return GP_OK;
}

int gp_set_option(int key, int value) {
    init_tuning();
    switch (key) {
        case GP_OPT_BWD_MODE: if (value < 0 || value > 2) return GP_ERR_SHAPE; g_tune.bwd_mode = value; return GP_OK;
        case GP_OPT_GIN_TILE_H: if (value < 1) return GP_ERR_SHAPE; g_tune.gin_th = value; return GP_OK;
        case GP_OPT_GIN_TILE_W: if (value < 1) return GP_ERR_SHAPE; g_tune.gin_tw = value; return GP_OK;
        case GP_OPT_GIN_THREADS: if (value != 128 && value != 192 && value != 256) return GP_ERR_SHAPE; g_tune.gin_nt = value; return GP_OK;
        case GP_OPT_FWD_MODE: if (value != 0 && value != 1) return GP_ERR_SHAPE; g_tune.fwd_mode = value; return GP_OK;
    }
    return GP_ERR_SHAPE;
}

int gp_get_option(int key) {
    init_tuning();
    switch (key) {
        case GP_OPT_BWD_MODE: return g_tune.bwd_mode;
        case GP_OPT_GIN_TILE_H: return g_tune.gin_th;
        case GP_OPT_GIN_TILE_W: return g_tune.gin_tw;
        case GP_OPT_GIN_THREADS: return g_tune.gin_nt;
        case GP_OPT_FWD_MODE: return g_tune.fwd_mode;
        case GP_OPT_LAST_FWD_KERNEL: return g_tune.last_fwd_kernel;
    }
    return GP_ERR_SHAPE;
}

int gp_dcnv3_forward(const void *input, const void *offset, const void *mask, void *out, const gp_dcnv3_desc *desc,
                     int dtype, void *stream) {
    return forward_impl<false>(input, offset, mask, out, desc, dtype, stream);
}

int gp_dcnv3_forward_softmax(const void *input, const void *offset, const void *mask_logits, void *out,
                             const gp_dcnv3_desc *desc, int dtype, void *stream) {
    return forward_impl<true>(input, offset, mask_logits, out, desc, dtype, stream);
}

int gp_dcnv3_forward_softmax_packed(const void *input, const void *offset_mask, void *out, long long pitch,
                                    const gp_dcnv3_desc *desc, int dtype, void *stream) {
    if (!offset_mask || !desc || pitch <= 0) return GP_ERR_NULL;
    const size_t es = elem_size(dtype);
    if (!es) return GP_ERR_DTYPE;
    const char *msk = (const char *)offset_mask + (size_t)desc->G * (desc->kh * desc->kw - desc->remove_center) * 2 * es;
    return forward_impl<true>(input, offset_mask, msk, out, desc, dtype, stream, pitch);
}

#endif
#if GP_PART != 1
size_t gp_dcnv3_backward_workspace(const gp_dcnv3_desc *desc, int dtype) {
    if (!desc || (dtype != GP_BF16 && dtype != GP_F16)) return 0;
    return (size_t)desc->N * desc->H * desc->W * desc->G * desc->gc * sizeof(float);
}

int gp_dcnv3_backward(const void *input, const void *offset, const void *mask, const void *grad_out, void *grad_input,
                      void *grad_offset, void *grad_mask, size_t grad_offset_elems, size_t grad_mask_elems,
                      void *workspace, size_t workspace_bytes, const gp_dcnv3_desc *desc, int dtype, void *stream) {
    if (!input || !offset || !mask || !grad_out || !grad_input || !grad_offset || !grad_mask) return GP_ERR_NULL;
    const size_t es = elem_size(dtype);
    if (!es) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(desc, p)) return e;
    if (!aligned16(input) || !aligned16(offset) || !aligned16(mask) || !aligned16(grad_out) || !aligned16(grad_input) ||
        !aligned16(grad_offset) || !aligned16(grad_mask) || !aligned16(workspace))
        return GP_ERR_ALIGN;
    const size_t n_in = (size_t)p.N * p.H * p.W * p.C;
    const size_t n_off = (size_t)p.n_units * p.P * 2, n_msk = (size_t)p.n_units * p.P;
    if (grad_offset_elems < n_off || grad_mask_elems < n_msk) return GP_ERR_SHAPE;
    const bool half16 = dtype == GP_BF16 || dtype == GP_F16;
    if (half16 && (!workspace || workspace_bytes < n_in * sizeof(float))) return GP_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;

    // zero the accumulation target and the rows of grad_offset/grad_mask beyond the flat prefix
    // (the reference zero-fills all three full tensors, dcnv3_cuda.cu:131-133; the prefix is fully overwritten here)
    void *acc = half16 ? workspace : grad_input;
    cudaError_t ce = cudaMemsetAsync(acc, 0, n_in * (half16 ? sizeof(float) : es), st);
    if (ce != cudaSuccess) return (int)ce;
    if (grad_offset_elems > n_off) {
        ce = cudaMemsetAsync((char *)grad_offset + n_off * es, 0, (grad_offset_elems - n_off) * es, st);
        if (ce != cudaSuccess) return (int)ce;
    }
    if (grad_mask_elems > n_msk) {
        ce = cudaMemsetAsync((char *)grad_mask + n_msk * es, 0, (grad_mask_elems - n_msk) * es, st);
        if (ce != cudaSuccess) return (int)ce;
    }

    int L = 0, vec = 0, Lg = 0, nt = 0;
    KParams pg;
    init_tuning();
    if (g_tune.bwd_mode == 2 && plan_bwd_fused(p, dtype, pg, &Lg)) {
        // one kernel: in-SM aggregated grad_input (sort + row walk) and the grad_offset / grad_mask gathers
        if (dtype == GP_F32) ce = launch_bwd_fused<float>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, pg, Lg, st);
        else if (dtype == GP_BF16) ce = launch_bwd_fused<__nv_bfloat16>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, pg, Lg, st);
        else ce = launch_bwd_fused<__half>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, pg, Lg, st);
        if (ce != cudaSuccess) return (int)ce;
    } else {
    const bool split = g_tune.bwd_mode == 1 && plan_gin_binned(p, dtype, pg, &Lg, &nt);
    if (plan_tiled(p, dtype, &L, &vec, !split)) {
        if (split) {
            // grad_offset / grad_mask (gathers + dot products), then grad_input (binned, no input reads)
            if (dtype == GP_F32) launch_bwd_tile<float, 4, false>(input, offset, mask, grad_out, nullptr, grad_offset, grad_mask, p, L, st);
            else if (dtype == GP_BF16 && vec == 8) launch_bwd_tile<__nv_bfloat16, 8, false>(input, offset, mask, grad_out, nullptr, grad_offset, grad_mask, p, L, st);
            else if (dtype == GP_BF16) launch_bwd_tile<__nv_bfloat16, 4, false>(input, offset, mask, grad_out, nullptr, grad_offset, grad_mask, p, L, st);
            else if (vec == 8) launch_bwd_tile<__half, 8, false>(input, offset, mask, grad_out, nullptr, grad_offset, grad_mask, p, L, st);
            else launch_bwd_tile<__half, 4, false>(input, offset, mask, grad_out, nullptr, grad_offset, grad_mask, p, L, st);
            ce = cudaGetLastError();
            if (ce != cudaSuccess) return (int)ce;
            if (dtype == GP_F32) ce = launch_gin_binned<float>(offset, mask, grad_out, (float *)acc, pg, Lg, nt, st);
            else if (dtype == GP_BF16) ce = launch_gin_binned<__nv_bfloat16>(offset, mask, grad_out, (float *)acc, pg, Lg, nt, st);
            else ce = launch_gin_binned<__half>(offset, mask, grad_out, (float *)acc, pg, Lg, nt, st);
            if (ce != cudaSuccess) return (int)ce;
        } else if (dtype == GP_F32) launch_bwd_tile<float, 4, true>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, p, L, st);
        else if (dtype == GP_BF16 && vec == 8) launch_bwd_tile<__nv_bfloat16, 8, true>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, p, L, st);
        else if (dtype == GP_BF16) launch_bwd_tile<__nv_bfloat16, 4, true>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, p, L, st);
        else if (vec == 8) launch_bwd_tile<__half, 8, true>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, p, L, st);
        else launch_bwd_tile<__half, 4, true>(input, offset, mask, grad_out, (float *)acc, grad_offset, grad_mask, p, L, st);
    } else {
        switch (dtype) {
            case GP_F32: launch_bwd_generic<float>(input, offset, mask, grad_out, acc, grad_offset, grad_mask, p, st); break;
            case GP_BF16: launch_bwd_generic<__nv_bfloat16>(input, offset, mask, grad_out, acc, grad_offset, grad_mask, p, st); break;
            case GP_F16: launch_bwd_generic<__half>(input, offset, mask, grad_out, acc, grad_offset, grad_mask, p, st); break;
            case GP_F64: launch_bwd_generic<double>(input, offset, mask, grad_out, acc, grad_offset, grad_mask, p, st); break;
        }
    }
    }
    ce = cudaGetLastError();
    if (ce != cudaSuccess) return (int)ce;
    if (half16) {
        const long long n = (long long)n_in;
        const unsigned grid = (unsigned)((n / 8 + 256) / 256);
        if (dtype == GP_BF16) cast_from_f32<__nv_bfloat16><<<grid, 256, 0, st>>>((const float *)acc, (__nv_bfloat16 *)grad_input, n);
        else cast_from_f32<__half><<<grid, 256, 0, st>>>((const float *)acc, (__half *)grad_input, n);
        count_launch();
        ce = cudaGetLastError();
    }
    return (int)ce;
}

#endif
#if GP_PART != 2
int gp_dcnv3_sample_index(const void *offset, int32_t *hw_low, uint8_t *flags, const gp_dcnv3_desc *desc, int dtype,
                          void *stream) {
    if (!offset || !hw_low || !flags) return GP_ERR_NULL;
    if (!elem_size(dtype)) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(desc, p)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((p.n_units + 255) / 256);
    switch (dtype) {
        case GP_F32: dcnv3_index_kernel<float><<<grid, 256, 0, st>>>((const float *)offset, hw_low, flags, p); break;
        case GP_BF16: dcnv3_index_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)offset, hw_low, flags, p); break;
        case GP_F16: dcnv3_index_kernel<__half><<<grid, 256, 0, st>>>((const __half *)offset, hw_low, flags, p); break;
        case GP_F64: dcnv3_index_kernel<double><<<grid, 256, 0, st>>>((const double *)offset, hw_low, flags, p); break;
    }
    count_launch();
    return (int)cudaGetLastError();
}

// ---- host-buffer entry points ---------------------------------------------------------------------------

namespace {
struct HostCache {
    void *buf[8] = {nullptr};
    size_t cap[8] = {0};
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined forward+backward call
    cudaEvent_t ev_up[64] = {nullptr}, ev_done[64] = {nullptr};
    int device = -1;
} g_hc;

cudaError_t hc_get(int slot, size_t bytes, void **out) {
    if (g_hc.cap[slot] < bytes) {
        if (g_hc.buf[slot]) cudaFree(g_hc.buf[slot]);
        g_hc.buf[slot] = nullptr;
        g_hc.cap[slot] = 0;
        cudaError_t e = cudaMalloc(&g_hc.buf[slot], bytes);
        if (e != cudaSuccess) return e;
        g_hc.cap[slot] = bytes;
    }
    *out = g_hc.buf[slot];
    return cudaSuccess;
}

// RAII: the *_host entry points run on `device` and hand the caller's current device back on every return path
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

cudaError_t hc_begin(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (g_hc.device != device) {
        gp_host_cache_release();
        g_hc.device = device;
    }
    if (!g_hc.stream) e = cudaStreamCreateWithFlags(&g_hc.stream, cudaStreamNonBlocking);
    return e;
}
}  // namespace

#define GP_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

int gp_host_cache_release(void) {
    for (int i = 0; i < 8; ++i) {
        if (g_hc.buf[i]) cudaFree(g_hc.buf[i]);
        g_hc.buf[i] = nullptr;
        g_hc.cap[i] = 0;
    }
    if (g_hc.stream) cudaStreamDestroy(g_hc.stream);
    if (g_hc.s_h2d) cudaStreamDestroy(g_hc.s_h2d);
    if (g_hc.s_d2h) cudaStreamDestroy(g_hc.s_d2h);
    for (int i = 0; i < 64; ++i) {
        if (g_hc.ev_up[i]) cudaEventDestroy(g_hc.ev_up[i]);
        if (g_hc.ev_done[i]) cudaEventDestroy(g_hc.ev_done[i]);
        g_hc.ev_up[i] = g_hc.ev_done[i] = nullptr;
    }
    g_hc.stream = g_hc.s_h2d = g_hc.s_d2h = nullptr;
    g_hc.device = -1;
    return GP_OK;
}

int gp_dcnv3_forward_host(const void *h_input, const void *h_offset, const void *h_mask, void *h_out,
                          size_t offset_elems, size_t mask_elems, const gp_dcnv3_desc *desc, int dtype, int device) {
    if (!h_input || !h_offset || !h_mask || !h_out) return GP_ERR_NULL;
    const size_t es = elem_size(dtype);
    if (!es) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(desc, p)) return e;
    const size_t n_in = (size_t)p.N * p.H * p.W * p.C, n_out = (size_t)p.N * p.Ho * p.Wo * p.C;
    const size_t n_off = (size_t)p.n_units * p.P * 2, n_msk = (size_t)p.n_units * p.P;
    if (offset_elems < n_off || mask_elems < n_msk) return GP_ERR_SHAPE;
    DeviceGuard guard_;
    GP_CUDA(hc_begin(device));
    void *d_in, *d_off, *d_msk, *d_out;
    GP_CUDA(hc_get(0, n_in * es, &d_in));
    GP_CUDA(hc_get(1, n_off * es, &d_off));   // only the flat prefix is ever read
    GP_CUDA(hc_get(2, n_msk * es, &d_msk));
    GP_CUDA(hc_get(3, n_out * es, &d_out));
    cudaStream_t st = g_hc.stream;
    GP_CUDA(cudaMemcpyAsync(d_in, h_input, n_in * es, cudaMemcpyHostToDevice, st));
    GP_CUDA(cudaMemcpyAsync(d_off, h_offset, n_off * es, cudaMemcpyHostToDevice, st));
    GP_CUDA(cudaMemcpyAsync(d_msk, h_mask, n_msk * es, cudaMemcpyHostToDevice, st));
    if (int e = gp_dcnv3_forward(d_in, d_off, d_msk, d_out, desc, dtype, st)) return e;
    GP_CUDA(cudaMemcpyAsync(h_out, d_out, n_out * es, cudaMemcpyDeviceToHost, st));
    GP_CUDA(cudaStreamSynchronize(st));
    return GP_OK;
}

int gp_dcnv3_backward_host(const void *h_input, const void *h_offset, const void *h_mask, const void *h_grad_out,
                           void *h_grad_input, void *h_grad_offset, void *h_grad_mask, size_t offset_elems,
                           size_t mask_elems, const gp_dcnv3_desc *desc, int dtype, int device) {
    if (!h_input || !h_offset || !h_mask || !h_grad_out || !h_grad_input || !h_grad_offset || !h_grad_mask)
        return GP_ERR_NULL;
    const size_t es = elem_size(dtype);
    if (!es) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(desc, p)) return e;
    const size_t n_in = (size_t)p.N * p.H * p.W * p.C, n_out = (size_t)p.N * p.Ho * p.Wo * p.C;
    const size_t n_off = (size_t)p.n_units * p.P * 2, n_msk = (size_t)p.n_units * p.P;
    if (offset_elems < n_off || mask_elems < n_msk) return GP_ERR_SHAPE;
    DeviceGuard guard_;
    GP_CUDA(hc_begin(device));
    void *d_in, *d_off, *d_msk, *d_go, *d_gi, *d_goff, *d_gmsk, *d_ws = nullptr;
    GP_CUDA(hc_get(0, n_in * es, &d_in));
    GP_CUDA(hc_get(1, n_off * es, &d_off));
    GP_CUDA(hc_get(2, n_msk * es, &d_msk));
    GP_CUDA(hc_get(3, n_out * es, &d_go));
    GP_CUDA(hc_get(4, n_in * es, &d_gi));
    GP_CUDA(hc_get(5, n_off * es, &d_goff));
    GP_CUDA(hc_get(6, n_msk * es, &d_gmsk));
    const size_t ws = gp_dcnv3_backward_workspace(desc, dtype);
    if (ws) GP_CUDA(hc_get(7, ws, &d_ws));
    cudaStream_t st = g_hc.stream;
    GP_CUDA(cudaMemcpyAsync(d_in, h_input, n_in * es, cudaMemcpyHostToDevice, st));
    GP_CUDA(cudaMemcpyAsync(d_off, h_offset, n_off * es, cudaMemcpyHostToDevice, st));
    GP_CUDA(cudaMemcpyAsync(d_msk, h_mask, n_msk * es, cudaMemcpyHostToDevice, st));
    GP_CUDA(cudaMemcpyAsync(d_go, h_grad_out, n_out * es, cudaMemcpyHostToDevice, st));
    if (int e = gp_dcnv3_backward(d_in, d_off, d_msk, d_go, d_gi, d_goff, d_gmsk, n_off, n_msk, d_ws, ws, desc, dtype, st))
        return e;
    GP_CUDA(cudaMemcpyAsync(h_grad_input, d_gi, n_in * es, cudaMemcpyDeviceToHost, st));
    GP_CUDA(cudaMemcpyAsync(h_grad_offset, d_goff, n_off * es, cudaMemcpyDeviceToHost, st));
    GP_CUDA(cudaMemcpyAsync(h_grad_mask, d_gmsk, n_msk * es, cudaMemcpyDeviceToHost, st));
    // rows beyond the flat prefix are zero by contract (dcnv3_cuda.cu:131-133): host-side fill, nothing to transfer
    if (offset_elems > n_off) memset((char *)h_grad_offset + n_off * es, 0, (offset_elems - n_off) * es);
    if (mask_elems > n_msk) memset((char *)h_grad_mask + n_msk * es, 0, (mask_elems - n_msk) * es);
    GP_CUDA(cudaStreamSynchronize(st));
    return GP_OK;
}

int gp_dcnv3_forward_backward_host(const void *h_input, const void *h_offset, const void *h_mask, const void *h_grad_out,
                                   void *h_out, void *h_grad_input, void *h_grad_offset, void *h_grad_mask,
                                   size_t offset_elems, size_t mask_elems, const gp_dcnv3_desc *desc, int dtype, int device,
                                   int chunks) {
    if (!h_input || !h_offset || !h_mask || !h_grad_out || !h_out || !h_grad_input || !h_grad_offset || !h_grad_mask)
        return GP_ERR_NULL;
    const size_t es = elem_size(dtype);
    if (!es) return GP_ERR_DTYPE;
    KParams p;
    if (int e = make_params(desc, p)) return e;
    const size_t img_in = (size_t)p.H * p.W * p.C, img_out = (size_t)p.Ho * p.Wo * p.C;
    const size_t img_off = (size_t)p.Ho * p.Wo * p.G * p.P * 2, img_msk = (size_t)p.Ho * p.Wo * p.G * p.P;   // flat prefix per RoI
    const size_t n_off = img_off * p.N, n_msk = img_msk * p.N;
    if (offset_elems < n_off || mask_elems < n_msk) return GP_ERR_SHAPE;
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    if (chunks > p.N) chunks = p.N;
    DeviceGuard guard_;
    GP_CUDA(hc_begin(device));
    if (!g_hc.s_h2d) GP_CUDA(cudaStreamCreateWithFlags(&g_hc.s_h2d, cudaStreamNonBlocking));
    if (!g_hc.s_d2h) GP_CUDA(cudaStreamCreateWithFlags(&g_hc.s_d2h, cudaStreamNonBlocking));
    for (int c = 0; c < chunks; ++c) {
        if (!g_hc.ev_up[c]) GP_CUDA(cudaEventCreateWithFlags(&g_hc.ev_up[c], cudaEventDisableTiming));
        if (!g_hc.ev_done[c]) GP_CUDA(cudaEventCreateWithFlags(&g_hc.ev_done[c], cudaEventDisableTiming));
    }
    char *d_in, *d_off, *d_msk, *d_go, *d_gi, *d_goff, *d_gmsk, *d_out;
    GP_CUDA(hc_get(0, img_in * p.N * es, (void **)&d_in));
    GP_CUDA(hc_get(1, n_off * es, (void **)&d_off));
    GP_CUDA(hc_get(2, n_msk * es, (void **)&d_msk));
    GP_CUDA(hc_get(3, img_out * p.N * es, (void **)&d_go));
    GP_CUDA(hc_get(4, img_in * p.N * es, (void **)&d_gi));
    GP_CUDA(hc_get(5, n_off * es, (void **)&d_goff));
    GP_CUDA(hc_get(6, n_msk * es, (void **)&d_gmsk));
    const bool half16 = dtype == GP_BF16 || dtype == GP_F16;
    // slot 7: [forward output | fp32 workspace of the 16-bit backward (one chunk)]
    int per = (p.N + chunks - 1) / chunks;
    // every chunk's device pointers must stay 16-byte aligned (gp_dcnv3_forward / backward require it): round the chunk
    // size up to the smallest RoI count whose offset / mask / image slices are multiples of 16 bytes (<= 16 RoIs)
    {
        int q = 1;
        while (q < 16 && (((size_t)q * img_msk * es) % 16 || ((size_t)q * img_off * es) % 16 || ((size_t)q * img_in * es) % 16 ||
                          ((size_t)q * img_out * es) % 16))
            ++q;
        per = ((per + q - 1) / q) * q;
        if (((size_t)per * img_msk * es) % 16 || ((size_t)per * img_in * es) % 16 || ((size_t)per * img_out * es) % 16) per = p.N;   // one chunk
    }
    const size_t ws_chunk = half16 ? img_in * per * sizeof(float) : 0;
    GP_CUDA(hc_get(7, img_out * p.N * es + ws_chunk + 256, (void **)&d_out));
    char *d_ws = half16 ? d_out + ((img_out * p.N * es + 255) / 256) * 256 : nullptr;
    const char *hi = (const char *)h_input, *ho = (const char *)h_offset, *hm = (const char *)h_mask, *hg = (const char *)h_grad_out;
    char *ro = (char *)h_out, *rgi = (char *)h_grad_input, *rgo = (char *)h_grad_offset, *rgm = (char *)h_grad_mask;
    // RoI chunks are independent: the flat offset/mask addressing is linear in the RoI index, so a chunk's rows start at
    // r0 * (Ho*Wo*G*P) of the same buffers.  H2D of chunk c+1, kernels of chunk c and D2H of chunk c-1 overlap (PCIe is full duplex).
    for (int c = 0, r0 = 0; c < chunks && r0 < p.N; ++c, r0 += per) {
        const int n = (r0 + per <= p.N) ? per : p.N - r0;
        gp_dcnv3_desc dc = *desc;
        dc.N = n;
        GP_CUDA(cudaMemcpyAsync(d_in + r0 * img_in * es, hi + r0 * img_in * es, n * img_in * es, cudaMemcpyHostToDevice, g_hc.s_h2d));
        GP_CUDA(cudaMemcpyAsync(d_off + r0 * img_off * es, ho + r0 * img_off * es, n * img_off * es, cudaMemcpyHostToDevice, g_hc.s_h2d));
        GP_CUDA(cudaMemcpyAsync(d_msk + r0 * img_msk * es, hm + r0 * img_msk * es, n * img_msk * es, cudaMemcpyHostToDevice, g_hc.s_h2d));
        GP_CUDA(cudaMemcpyAsync(d_go + r0 * img_out * es, hg + r0 * img_out * es, n * img_out * es, cudaMemcpyHostToDevice, g_hc.s_h2d));
        GP_CUDA(cudaEventRecord(g_hc.ev_up[c], g_hc.s_h2d));
        GP_CUDA(cudaStreamWaitEvent(g_hc.stream, g_hc.ev_up[c], 0));
        if (int e = gp_dcnv3_forward(d_in + r0 * img_in * es, d_off + r0 * img_off * es, d_msk + r0 * img_msk * es,
                                     d_out + r0 * img_out * es, &dc, dtype, g_hc.stream)) return e;
        if (int e = gp_dcnv3_backward(d_in + r0 * img_in * es, d_off + r0 * img_off * es, d_msk + r0 * img_msk * es,
                                      d_go + r0 * img_out * es, d_gi + r0 * img_in * es, d_goff + r0 * img_off * es,
                                      d_gmsk + r0 * img_msk * es, n * img_off, n * img_msk, d_ws, ws_chunk, &dc, dtype, g_hc.stream)) return e;
        GP_CUDA(cudaEventRecord(g_hc.ev_done[c], g_hc.stream));
        GP_CUDA(cudaStreamWaitEvent(g_hc.s_d2h, g_hc.ev_done[c], 0));
        GP_CUDA(cudaMemcpyAsync(ro + r0 * img_out * es, d_out + r0 * img_out * es, n * img_out * es, cudaMemcpyDeviceToHost, g_hc.s_d2h));
        GP_CUDA(cudaMemcpyAsync(rgi + r0 * img_in * es, d_gi + r0 * img_in * es, n * img_in * es, cudaMemcpyDeviceToHost, g_hc.s_d2h));
        GP_CUDA(cudaMemcpyAsync(rgo + r0 * img_off * es, d_goff + r0 * img_off * es, n * img_off * es, cudaMemcpyDeviceToHost, g_hc.s_d2h));
        GP_CUDA(cudaMemcpyAsync(rgm + r0 * img_msk * es, d_gmsk + r0 * img_msk * es, n * img_msk * es, cudaMemcpyDeviceToHost, g_hc.s_d2h));
    }
    // rows beyond the flat prefix are zero by contract (dcnv3_cuda.cu:131-133): host-side fill, nothing to transfer
    if (offset_elems > n_off) memset(rgo + n_off * es, 0, (offset_elems - n_off) * es);
    if (mask_elems > n_msk) memset(rgm + n_msk * es, 0, (mask_elems - n_msk) * es);
    GP_CUDA(cudaStreamSynchronize(g_hc.s_d2h));
    return GP_OK;
}

#endif
}  // extern "C"
