/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference DCNv3 CUDA kernels.
 * Never imported, linked or executed by the product path (givepose_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
 *
 * This header is included twice by dcnv3_oracle.c with
 *   REAL  = float / double         (the reference's opmath_t)
 *   SUFFIX = f32 / f64
 * and restates, in the reference's operation order:
 *   forward   network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:216-282  (dcnv3_im2col_gpu_kernel)
 *             + :32-80 (dcnv3_im2col_bilinear)
 *   backward  :386-487 (dcnv3_col2im_gpu_kernel_shm_blocksize_aware_reduce_v2; all six backward
 *             variants compute the same sums, they differ only in how the per-group reduction is done)
 *             + :82-147 (dcnv3_col2im_bilinear)
 *   host      network/ops_dcnv3/src/cuda/dcnv3_cuda.cu:21-85, :87-174 (output size, flat
 *             offset/mask addressing, im2col_step chunking -- chunking does not change any address:
 *             chunk base n*step*Ho*Wo*G*P(*2) + in-chunk sampling_index*P == global (q*G+g)*P).
 *
 * Floating-point contraction: the reference is built by nvcc with its default -fmad=true, so
 *   p0_w_ = p0_w - c*scale        and      loc_w = p0_w_ + (i*dil + off)*scale
 * are single-rounding FMAs in the reference binary.  We write them as explicit FMA() so that the
 * floor()/bounds decisions (the "bit-exact" part of the contract) do not depend on this file's
 * compiler flags (it is compiled with -ffp-contract=off).  For the power-of-two offset_scale values
 * used everywhere in the reference (1.0, 2.0) FMA and mul+add are identical anyway.
 */

#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* one sampling point: location, integer corners, validity (dcnv3_im2col_cuda.cuh:249-269, :39-75) */
typedef struct {
    REAL loc_h, loc_w;
    int in_range;           /* loc_h > -1 && loc_w > -1 && loc_h < H && loc_w < W      (:268-269) */
    int h_low, w_low;       /* floor()                                                 (:39-40)  */
    int ok1, ok2, ok3, ok4; /* per-corner bounds checks                                (:57,62,67,72) */
    REAL lh, lw, hh, hw;
} FN(gpo_point);

static inline void FN(gpo_locate)(FN(gpo_point) * pt, REAL p0_h_, REAL p0_w_, int i, int j, int dil_h,
                                  int dil_w, REAL off_w, REAL off_h, REAL scale, int H, int W) {
    /* :263-266  loc = p0_ + (i*dil + off) * scale  (FMA, see header comment) */
    pt->loc_w = FMA((REAL)(i * dil_w) + off_w, scale, p0_w_);
    pt->loc_h = FMA((REAL)(j * dil_h) + off_h, scale, p0_h_);
    pt->in_range = (pt->loc_h > (REAL)-1 && pt->loc_w > (REAL)-1 && pt->loc_h < (REAL)H && pt->loc_w < (REAL)W);
    pt->h_low = 0; pt->w_low = 0; pt->ok1 = pt->ok2 = pt->ok3 = pt->ok4 = 0;
    pt->lh = pt->lw = pt->hh = pt->hw = 0;
    if (!pt->in_range) return;
    pt->h_low = (int)FLOOR(pt->loc_h);
    pt->w_low = (int)FLOOR(pt->loc_w);
    const int h_high = pt->h_low + 1, w_high = pt->w_low + 1;
    pt->lh = pt->loc_h - (REAL)pt->h_low;
    pt->lw = pt->loc_w - (REAL)pt->w_low;
    pt->hh = (REAL)1 - pt->lh;
    pt->hw = (REAL)1 - pt->lw;
    pt->ok1 = (pt->h_low >= 0 && pt->w_low >= 0);
    pt->ok2 = (pt->h_low >= 0 && w_high <= W - 1);
    pt->ok3 = (h_high <= H - 1 && pt->w_low >= 0);
    pt->ok4 = (h_high <= H - 1 && w_high <= W - 1);
}

/*
 * Forward.  in (N,H,W,G*gc); off: flat array, row (q*G+g) holds P*(w,h) pairs; mask: flat, row
 * (q*G+g) holds P weights; out (N,Ho,Wo,G*gc); q = (b*Ho+oh)*Wo+ow.   off/mask may be LARGER than
 * N*Ho*Wo*G*P*(2) elements (the stride-2 quirk): only the flat prefix is read.
 */
void FN(gpo_dcnv3_forward)(const REAL *in, const REAL *off, const REAL *mask, REAL *out, int N, int H, int W,
                           int G, int gc, int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                           REAL scale, int remove_center, int Ho, int Wo) {
    const int C = G * gc;
    const int P = kh * kw - remove_center;
    const int center_h = kh / 2, center_w = kw / 2;
    const long long npix = (long long)N * Ho * Wo;
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < npix; ++q) {
        const int ow = (int)(q % Wo), oh = (int)((q / Wo) % Ho), b = (int)(q / ((long long)Wo * Ho));
        const int p0_w = ((dw * (kw - 1)) >> 1) - pw + ow * sw;     /* :232-233 */
        const int p0_h = ((dh * (kh - 1)) >> 1) - ph + oh * sh;     /* :235-236 */
        const REAL p0_w_ = FMA(-(REAL)((dw * (kw - 1)) >> 1), scale, (REAL)p0_w); /* :249-250 */
        const REAL p0_h_ = FMA(-(REAL)((dh * (kh - 1)) >> 1), scale, (REAL)p0_h); /* :251-252 */
        const REAL *im = in + (long long)b * H * W * C;
        for (int g = 0; g < G; ++g) {
            REAL *o = out + q * C + (long long)g * gc;
            for (int c = 0; c < gc; ++c) o[c] = 0;
            long long wptr = (q * G + g) * P;   /* data_weight_ptr :243 */
            long long lptr = wptr * 2;          /* data_loc_w_ptr  :244 */
            for (int i = 0; i < kw; ++i) {      /* width is the OUTER loop :257 */
                for (int j = 0; j < kh; ++j) {
                    if (i != center_w || j != center_h || !remove_center) {
                        FN(gpo_point) pt;
                        FN(gpo_locate)(&pt, p0_h_, p0_w_, i, j, dh, dw, off[lptr], off[lptr + 1], scale, H, W);
                        const REAL wgt = mask[wptr];
                        if (pt.in_range) {
                            const long long r_lo = ((long long)pt.h_low * W) * C, r_hi = r_lo + (long long)W * C;
                            const long long c_lo = (long long)pt.w_low * C, c_hi = c_lo + C;
                            const REAL w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                            for (int c = 0; c < gc; ++c) {
                                const int ch = g * gc + c;
                                const REAL v1 = pt.ok1 ? im[r_lo + c_lo + ch] : 0;
                                const REAL v2 = pt.ok2 ? im[r_lo + c_hi + ch] : 0;
                                const REAL v3 = pt.ok3 ? im[r_hi + c_lo + ch] : 0;
                                const REAL v4 = pt.ok4 ? im[r_hi + c_hi + ch] : 0;
                                const REAL val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);   /* :78 */
                                o[c] += val * wgt;                                         /* :270-273 */
                            }
                        }
                        wptr += 1;
                        lptr += 2;
                    }
                }
            }
        }
    }
}

/*
 * Index / bounds restatement: for every (q, g, p) writes
 *   hw_low[2*k+0] = h_low, hw_low[2*k+1] = w_low   (0 when the sample is out of range)
 *   flags[k] = bit0 in_range | bit1 ok1 | bit2 ok2 | bit3 ok3 | bit4 ok4
 * with k = (q*G+g)*P + p.  This is the integer part of the contract (bit-exact).
 */
void FN(gpo_dcnv3_index)(const REAL *off, int *hw_low, unsigned char *flags, int N, int H, int W, int G, int kh,
                         int kw, int sh, int sw, int ph, int pw, int dh, int dw, REAL scale, int remove_center,
                         int Ho, int Wo) {
    const int P = kh * kw - remove_center;
    const int center_h = kh / 2, center_w = kw / 2;
    const long long npix = (long long)N * Ho * Wo;
#pragma omp parallel for schedule(static)
    for (long long q = 0; q < npix; ++q) {
        const int ow = (int)(q % Wo), oh = (int)((q / Wo) % Ho);
        const int p0_w = ((dw * (kw - 1)) >> 1) - pw + ow * sw;
        const int p0_h = ((dh * (kh - 1)) >> 1) - ph + oh * sh;
        const REAL p0_w_ = FMA(-(REAL)((dw * (kw - 1)) >> 1), scale, (REAL)p0_w);
        const REAL p0_h_ = FMA(-(REAL)((dh * (kh - 1)) >> 1), scale, (REAL)p0_h);
        for (int g = 0; g < G; ++g) {
            long long k = (q * G + g) * P;
            for (int i = 0; i < kw; ++i)
                for (int j = 0; j < kh; ++j)
                    if (i != center_w || j != center_h || !remove_center) {
                        FN(gpo_point) pt;
                        FN(gpo_locate)(&pt, p0_h_, p0_w_, i, j, dh, dw, off[2 * k], off[2 * k + 1], scale, H, W);
                        hw_low[2 * k] = pt.h_low;
                        hw_low[2 * k + 1] = pt.w_low;
                        flags[k] = (unsigned char)(pt.in_range | (pt.ok1 << 1) | (pt.ok2 << 2) | (pt.ok3 << 3) |
                                                   (pt.ok4 << 4));
                        ++k;
                    }
        }
    }
}

/*
 * Backward.  grad_in (N,H,W,C), grad_off / grad_mask: flat prefix of N*Ho*Wo*G*P(*2) elements is
 * WRITTEN (the caller zero-fills the full-size buffers first, as dcnv3_cuda.cu:131-133 does).
 * grad_in must be zero on entry; it is accumulated image by image (deterministic order here; the
 * reference uses atomicAdd, :116-140, so its order is not defined).
 */
void FN(gpo_dcnv3_backward)(const REAL *in, const REAL *off, const REAL *mask, const REAL *grad_out,
                            REAL *grad_in, REAL *grad_off, REAL *grad_mask, int N, int H, int W, int G, int gc,
                            int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, REAL scale,
                            int remove_center, int Ho, int Wo) {
    const int C = G * gc;
    const int P = kh * kw - remove_center;
    const int center_h = kh / 2, center_w = kw / 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < N; ++b) {
        const REAL *im = in + (long long)b * H * W * C;
        REAL *gim = grad_in + (long long)b * H * W * C;
        for (int oh = 0; oh < Ho; ++oh)
            for (int ow = 0; ow < Wo; ++ow) {
                const long long q = ((long long)b * Ho + oh) * Wo + ow;
                const int p0_w = ((dw * (kw - 1)) >> 1) - pw + ow * sw;
                const int p0_h = ((dh * (kh - 1)) >> 1) - ph + oh * sh;
                const REAL p0_w_ = FMA(-(REAL)((dw * (kw - 1)) >> 1), scale, (REAL)p0_w);
                const REAL p0_h_ = FMA(-(REAL)((dh * (kh - 1)) >> 1), scale, (REAL)p0_h);
                for (int g = 0; g < G; ++g) {
                    const REAL *go = grad_out + q * C + (long long)g * gc;
                    long long wptr = (q * G + g) * P;
                    long long lptr = wptr * 2;
                    for (int i = 0; i < kw; ++i)
                        for (int j = 0; j < kh; ++j)
                            if (i != center_w || j != center_h || !remove_center) {
                                FN(gpo_point) pt;
                                FN(gpo_locate)(&pt, p0_h_, p0_w_, i, j, dh, dw, off[lptr], off[lptr + 1], scale, H, W);
                                const REAL wgt = mask[wptr];
                                REAL s_mask = 0, s_ow = 0, s_oh = 0;   /* sums over the gc channels, :458-476 */
                                if (pt.in_range) {
                                    const long long r_lo = ((long long)pt.h_low * W) * C, r_hi = r_lo + (long long)W * C;
                                    const long long c_lo = (long long)pt.w_low * C, c_hi = c_lo + C;
                                    const REAL w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw,
                                               w4 = pt.lh * pt.lw;
                                    for (int c = 0; c < gc; ++c) {
                                        const int ch = g * gc + c;
                                        const REAL top_grad = go[c];
                                        const REAL top_grad_im = top_grad * wgt;        /* :107 */
                                        REAL gh = 0, gw = 0, v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                                        if (pt.ok1) { v1 = im[r_lo + c_lo + ch]; gh -= pt.hw * v1; gw -= pt.hh * v1; gim[r_lo + c_lo + ch] += w1 * top_grad_im; }
                                        if (pt.ok2) { v2 = im[r_lo + c_hi + ch]; gh -= pt.lw * v2; gw += pt.hh * v2; gim[r_lo + c_hi + ch] += w2 * top_grad_im; }
                                        if (pt.ok3) { v3 = im[r_hi + c_lo + ch]; gh += pt.hw * v3; gw -= pt.lh * v3; gim[r_hi + c_lo + ch] += w3 * top_grad_im; }
                                        if (pt.ok4) { v4 = im[r_hi + c_hi + ch]; gh += pt.lw * v4; gw += pt.lh * v4; gim[r_hi + c_hi + ch] += w4 * top_grad_im; }
                                        const REAL val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                                        s_mask += top_grad * val;                       /* :144 */
                                        s_ow += scale * gw * top_grad_im;               /* :145 */
                                        s_oh += scale * gh * top_grad_im;               /* :146 */
                                    }
                                }
                                grad_mask[wptr] = s_mask;
                                grad_off[lptr] = s_ow;
                                grad_off[lptr + 1] = s_oh;
                                wptr += 1;
                                lptr += 2;
                            }
                }
            }
    }
}

#undef CAT_
#undef CAT
#undef FN
