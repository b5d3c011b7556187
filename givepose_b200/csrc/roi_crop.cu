// givepose_b200 -- RoI input pipeline in front of PoseNet.forward: the per-detection crops the
// reference makes on the host with OpenCV, one RoI at a time (evaluation/load_data_eval.py:256-289,
// datasets/load_data_nocs.py:277-305, tools/dataset_utils.py:8-30,101-157), done for the whole batch on the device.
//
//   reference, per RoI                                              here
//   trans = get_affine_transform(center, scale, 0, out)            gp_roi_affine_inverse (host, double; restates the float32
//     -> cv2.getAffineTransform (6x6 LU solve in double)              point construction and OpenCV's LU bit for bit) followed by
//   cv2.warpAffine(img, trans, (out, out), INTER_NEAREST)            the inversion cv::warpAffine does before sampling
//     -> fixed-point source indices, BORDER_CONSTANT 0             roi_crop_kernel: the same 10-bit fixed-point index arithmetic
//   (roi / 255.0 - mean) / std, HWC -> CHW, float64 -> float32      a 3 x 256 entry table (computed in double on the host)
//   get_2d_coord_np(W, H) cropped the same way at out_res           evaluated from the source index, never materialised
//   mask.astype(float32) / (mask == inst_id) cropped at img_size    same gather on a uint8 mask plane
//
// Everything here is integer / index work: results are bit-exact against OpenCV 4.8 (the reference's pin, GIVEPose_env.yml:250)
// and against the cv2 in this image (tests/test_roi_gpu.py).  HBM-bound: 1 MB of fp32 crops written per RoI.
#include <cmath>
#include <cstdint>

#include "gp_common.cuh"
#include "givepose_b200.h"

namespace gp {

// cv::warpAffine, INTER_NEAREST (imgwarp.cpp, OpenCV 4.8: hal::warpAffine + WarpAffineInvoker):
//   adelta[x] = cvRound(M0*x*1024), bdelta[x] = cvRound(M3*x*1024)
//   X0 = cvRound((M1*y + M2)*1024) + 512, Y0 = cvRound((M4*y + M5)*1024) + 512
//   X = saturate_cast<short>((X0 + adelta[x]) >> 10), Y likewise; remap(INTER_NEAREST, BORDER_CONSTANT, 0)
// cvRound is round-half-to-even (cvtsd2si); the products are formed exactly as written (no contraction).

// grid (ceil(S*S/4 / 256), B, 2): z == 0 -> the S x S crops (roi_img, roi_mask), z == 1 -> the R x R coordinate crop.
// One thread = 4 consecutive output pixels of a row (S, R multiples of 4): the row terms X0 / Y0 are computed once, every
// plane is written with 16-byte streaming stores.
__global__ void __launch_bounds__(256)
roi_crop_kernel(const uint8_t *__restrict__ images, int H, int W, const int *__restrict__ image_index,
                const uint8_t *__restrict__ masks, const int *__restrict__ mask_index, const int *__restrict__ inst_id,
                const double *__restrict__ minv_img, const double *__restrict__ minv_out, const float *__restrict__ lut,
                float *__restrict__ roi_img, float *__restrict__ roi_mask, float *__restrict__ roi_coord, int S, int R) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;   // quad index
    const bool coord = blockIdx.z != 0;
    const int n = coord ? R : S;
    if (q >= n * n / 4 || (coord && !roi_coord)) return;
    const int y = q / (n / 4), x0 = 4 * (q - y * (n / 4));
    const double *m = (coord ? minv_out : minv_img) + 6 * b;
    const double m0 = m[0], m3 = m[3], dy = (double)y;
    const int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[1], dy), m[2]), 1024.0)) + 512;
    const int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(m[4], dy), m[5]), 1024.0)) + 512;
    int X[4], Y[4];
    bool in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double dx = (double)(x0 + k);
        const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(m0, dx), 1024.0));
        const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(m3, dx), 1024.0));
        X[k] = max(-32768, min(32767, (X0 + adelta) >> 10));
        Y[k] = max(-32768, min(32767, (Y0 + bdelta) >> 10));
        in[k] = (unsigned)X[k] < (unsigned)W && (unsigned)Y[k] < (unsigned)H;
    }
    const long long p = (long long)y * n + x0;
    if (!coord) {
        if (roi_img) {
            const uint8_t *img = images + (long long)image_index[b] * H * W * 3;
            float r[4], g[4], bl[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // border pixels are 0 BEFORE the normalisation (load_data_eval.py:272-279)
                unsigned cr = 0, cg = 0, cb = 0;
                if (in[k]) {
                    const uint8_t *px = img + ((long long)Y[k] * W + X[k]) * 3;
                    cr = px[0]; cg = px[1]; cb = px[2];
                }
                r[k] = __ldg(lut + cr); g[k] = __ldg(lut + 256 + cg); bl[k] = __ldg(lut + 512 + cb);
            }
            float *o = roi_img + (long long)b * 3 * S * S + p;
            __stcs(reinterpret_cast<float4 *>(o), make_float4(r[0], r[1], r[2], r[3]));
            __stcs(reinterpret_cast<float4 *>(o + (long long)S * S), make_float4(g[0], g[1], g[2], g[3]));
            __stcs(reinterpret_cast<float4 *>(o + 2ll * S * S), make_float4(bl[0], bl[1], bl[2], bl[3]));
        }
        if (roi_mask) {
            const uint8_t *mk = masks + (long long)mask_index[b] * H * W;
            const int id = inst_id[b];
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = 0.f;
                if (in[k]) {
                    const unsigned mv = mk[(long long)Y[k] * W + X[k]];
                    v[k] = id < 0 ? (float)mv : (mv == (unsigned)id ? 1.f : 0.f);   // mask.astype(float32) | (mask == inst_id)
                }
            }
            __stcs(reinterpret_cast<float4 *>(roi_mask + (long long)b * S * S + p), make_float4(v[0], v[1], v[2], v[3]));
        }
    } else {
        // get_2d_coord_np (tools/dataset_utils.py:8-30): float32 x, float32 scalar (n-1)/2: fl((x - c) / c)
        const float cw = (float)(((double)W - 1.0) / 2.0), ch = (float)(((double)H - 1.0) / 2.0);
        float vx[4], vy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            vx[k] = in[k] ? __fdiv_rn(__fsub_rn((float)X[k], cw), cw) : 0.f;
            vy[k] = in[k] ? __fdiv_rn(__fsub_rn((float)Y[k], ch), ch) : 0.f;
        }
        float *o = roi_coord + (long long)b * 2 * R * R + p;
        __stcs(reinterpret_cast<float4 *>(o), make_float4(vx[0], vx[1], vx[2], vx[3]));
        __stcs(reinterpret_cast<float4 *>(o + (long long)R * R), make_float4(vy[0], vy[1], vy[2], vy[3]));
    }
}

// cv2.resize(image, (dw, dh)) -- INTER_LINEAR on 8-bit pixels (load_data_eval.py:336, FLAGS.resize_full) -- followed by the
// loaders' normalisation.  OpenCV's 8-bit linear resize is fixed point (resize.cpp, INTER_RESIZE_COEF_BITS = 11):
//   fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx, clamped at the borders (fx = 0)
//   alpha = { cvRound((1 - fx) * 2048), cvRound(fx * 2048) }  (float arithmetic, saturate_cast<short>), beta likewise for rows
//   row[k] = S[sx] * alpha0 + S[sx + 1] * alpha1                                            (HResizeLinear, int)
//   dst = (((beta0 * (row0 >> 4)) >> 16) + ((beta1 * (row1 >> 4)) >> 16) + 2) >> 2          (VResizeLinear<uchar>)
// with scale_x = 1.0 / ((double)dw / W).  Bit-exact against cv2 for source sizes >= the output size (the frames of the loaders
// are 480 x 640 -> 256 x 256); OpenCV takes a different path when upscaling (differences of 1 in ~0.1 % of the pixels), so the
// host refuses H < dh or W < dw.  One thread per output pixel, three channels.
__device__ __forceinline__ void resize_coeff(int d, double scale, int n_src, int &s, int &a0, int &a1) {
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n_src - 1) { f = 0.f; s = n_src - 1; }
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

__global__ void __launch_bounds__(256)
resize_linear_u8_norm_kernel(const uint8_t *__restrict__ images, const int *__restrict__ image_index, const float *__restrict__ lut,
                             float *__restrict__ out, int H, int W, int dh, int dw, double scale_y, double scale_x) {
    const int b = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dh * dw) return;
    const int dy = p / dw, dx = p - dy * dw;
    int sx, sy, a0, a1, b0, b1;
    resize_coeff(dx, scale_x, W, sx, a0, a1);
    resize_coeff(dy, scale_y, H, sy, b0, b1);
    const int sx1 = min(sx + 1, W - 1), sy1 = min(sy + 1, H - 1);
    const uint8_t *img = images + (long long)(image_index ? image_index[b] : b) * H * W * 3;
    const uint8_t *r0 = img + (long long)sy * W * 3, *r1 = img + (long long)sy1 * W * 3;
    float *o = out + (long long)b * 3 * dh * dw + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int row0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
        const int row1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
        int v = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2;
        v = max(0, min(255, v));
        __stcs(o + (long long)c * dh * dw, __ldg(lut + 256 * c + v));
    }
}

// ---- host: the affine the reference builds per RoI, in double, bit for bit -------------------------------------------
// cv::LU (matrix_decomp.cpp LUImpl<double>): partial pivoting, elimination with alpha = A[j][i] * (-1/A[i][i]).
static bool lu_solve6(double A[6][6], double b[6]) {
    const int m = 6;
    for (int i = 0; i < m; ++i) {
        int k = i;
        for (int j = i + 1; j < m; ++j)
            if (std::fabs(A[j][i]) > std::fabs(A[k][i])) k = j;
        if (std::fabs(A[k][i]) < 2.220446049250313e-16 * 100) return false;   // DBL_EPSILON*100 (cv::LU)
        if (k != i) {
            for (int j = i; j < m; ++j) { const double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
            const double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        const double d = -1 / A[i][i];
        for (int j = i + 1; j < m; ++j) {
            const double alpha = A[j][i] * d;
            for (int kk = i + 1; kk < m; ++kk) A[j][kk] += alpha * A[i][kk];   // built with -ffp-contract=off: product rounded, then added
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int k = i + 1; k < m; ++k) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    return true;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_roi_affine_inverse(const double *center, const double *scale, int B, int out_size, double *minv) {
    if (!center || !scale || !minv) return GP_ERR_NULL;
    if (B < 0 || out_size <= 0) return GP_ERR_SHAPE;
    for (int r = 0; r < B; ++r) {
        // get_affine_transform(center, (scale, scale), rot = 0, (out, out)), tools/dataset_utils.py:116-157, as reached through
        // crop_resize_by_warp_affine (:106-110): scale arrives as a tuple of Python floats, so src_w * -0.5 and the sums are
        // double; the three points are then STORED in float32 arrays (:142-150)
        const double cx = center[2 * r], cy = center[2 * r + 1], sc = scale[r];
        const double zero = sc * 0.0;                           // scale_tmp * shift with shift = (0, 0)
        const double dir_x = 0.0 * 1.0 - (sc * -0.5) * 0.0;     // get_dir([0, src_w * -0.5], 0): sn = 0, cs = 1 (:159-166)
        const double dir_y = 0.0 * 0.0 + (sc * -0.5) * 1.0;
        float src[3][2], dst[3][2];
        src[0][0] = (float)(cx + zero);
        src[0][1] = (float)(cy + zero);
        src[1][0] = (float)((cx + dir_x) + zero);
        src[1][1] = (float)((cy + dir_y) + zero);
        const float half = (float)((double)out_size * 0.5);
        dst[0][0] = half; dst[0][1] = half;
        dst[1][0] = half + 0.f;                                 // np.array([w/2, h/2], float32) + dst_dir (float32)
        dst[1][1] = half + (float)((double)out_size * -0.5);
        // get_3rd_point(a, b): direct = a - b; b + [-direct[1], direct[0]]   (float32 arithmetic)
        auto third = [](const float a[2], const float b[2], float o[2]) {
            const float d0 = a[0] - b[0], d1 = a[1] - b[1];
            o[0] = b[0] + (-d1);
            o[1] = b[1] + d0;
        };
        third(src[0], src[1], src[2]);
        third(dst[0], dst[1], dst[2]);
        // cv::getAffineTransform(src, dst): 6x6 system in double, cv::solve(DECOMP_LU)
        double A[6][6] = {{0}}, bb[6];
        for (int i = 0; i < 3; ++i) {
            A[2 * i][0] = A[2 * i + 1][3] = src[i][0];
            A[2 * i][1] = A[2 * i + 1][4] = src[i][1];
            A[2 * i][2] = A[2 * i + 1][5] = 1;
            bb[2 * i] = dst[i][0];
            bb[2 * i + 1] = dst[i][1];
        }
        if (!lu_solve6(A, bb)) return GP_ERR_SHAPE;            // degenerate RoI (scale 0)
        // cv::warpAffine without WARP_INVERSE_MAP inverts the 2x3 matrix first (imgwarp.cpp)
        double M[6] = {bb[0], bb[1], bb[2], bb[3], bb[4], bb[5]};
        double D = M[0] * M[4] - M[1] * M[3];
        D = D != 0 ? 1. / D : 0;
        const double A11 = M[4] * D, A22 = M[0] * D;
        M[0] = A11; M[1] *= -D;
        M[3] *= -D; M[4] = A22;
        const double b1 = -M[0] * M[2] - M[1] * M[5];
        const double b2 = -M[3] * M[2] - M[4] * M[5];
        M[2] = b1; M[5] = b2;
        for (int i = 0; i < 6; ++i) minv[6 * r + i] = M[i];
    }
    return GP_OK;
}

int gp_roi_crop(const uint8_t *images, int n_images, int H, int W, const int *image_index, const uint8_t *masks, int n_masks,
                const int *mask_index, const int *inst_id, const double *minv_img, const double *minv_out, const float *lut,
                float *roi_img, float *roi_mask, float *roi_coord_2d, int B, int img_size, int out_res, void *stream) {
    if (!minv_img || !minv_out) return GP_ERR_NULL;
    if (roi_img && (!images || !image_index || !lut)) return GP_ERR_NULL;
    if (roi_mask && (!masks || !mask_index || !inst_id)) return GP_ERR_NULL;
    if (B < 0 || B > 65535 || H <= 0 || W <= 0 || H > 32767 || W > 32767 || img_size <= 0 || out_res <= 0 || n_images < 0 || n_masks < 0)
        return GP_ERR_SHAPE;
    if (img_size % 4 || out_res % 4) return GP_ERR_UNSUPPORTED;   // 16-byte stores (the loaders use 256 / 64)
    if (B == 0) return GP_OK;
    const int big = img_size > out_res ? img_size : out_res;
    const dim3 grid((unsigned)((big * big / 4 + 255) / 256), (unsigned)B, roi_coord_2d ? 2u : 1u);
    roi_crop_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, H, W, image_index, masks, mask_index, inst_id, minv_img, minv_out, lut,
                                                            roi_img, roi_mask, roi_coord_2d, img_size, out_res);
    count_launch();
    return (int)cudaGetLastError();
}

int gp_resize_linear_u8_normalize(const uint8_t *images, int n_images, int H, int W, const int *image_index, const float *lut, float *out,
                                  int B, int dh, int dw, void *stream) {
    if (!images || !lut || !out) return GP_ERR_NULL;
    if (B < 0 || B > 65535 || n_images <= 0 || H <= 0 || W <= 0 || dh <= 0 || dw <= 0) return GP_ERR_SHAPE;
    if (H < dh || W < dw) return GP_ERR_UNSUPPORTED;   // upscaling: OpenCV takes another path (see the kernel comment)
    if (B == 0) return GP_OK;
    const double scale_x = 1.0 / ((double)dw / (double)W), scale_y = 1.0 / ((double)dh / (double)H);   // cv::resize: 1. / inv_scale
    const dim3 grid((unsigned)((dh * dw + 255) / 256), (unsigned)B);
    resize_linear_u8_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, image_index, lut, out, H, W, dh, dw, scale_y, scale_x);
    count_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
