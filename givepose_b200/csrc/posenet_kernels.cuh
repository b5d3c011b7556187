// givepose_b200 -- fused glue kernels of the per-RoI PoseNet forward around the DCNv3 core (sm_100a).
//
//   dwconv_ln_gelu   modules/dcnv3.py:269-283,329  x1 = GELU(LayerNorm_{eps}(DWConv3x3(x)))  channel-last, evaluated
//                    only for the first `rows` pixels of the flat (N*H*W) pixel list: the sampler reads offset/mask
//                    through their flat [N*Ho*Wo] prefix, so at stride 2 three quarters of x1 are dead.
//   gn_stats/apply   GroupNorm(32) + ReLU / exact GELU on channel-last activations (get_norm "GN", layer_utils.py:32-60;
//                    conv_module.py order conv -> norm -> act), optionally fused with the align_corners=True bilinear x2
//                    upsampling that follows it in TopDownXyzHead (xyz_head.py:262-265).
//   pose_decode      fc_r/fc_t/fc_z outputs -> rot6d_to_mat_batch (rot_reps.py:34-55) -> back-projection
//                    (pose_from_pred_centroid_z.py:78-119) -> allo->ego rotation (pose_utils/utils.py:29-60), one
//                    thread per RoI on the device instead of the reference's per-RoI host loop (:139-157).
#pragma once

#include "gp_common.cuh"

namespace gp {

// ---------------------------------------------------------------------------------------------------
// depthwise 3x3 (pad 1, stride 1) + bias -> LayerNorm over C -> exact GELU, channel-last, fp32 math.
// One warp per pixel; lane l owns channels {4*(l + 32*j) .. +3}, j < C/128 (coalesced 512-byte rows).
// ---------------------------------------------------------------------------------------------------
template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
    static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) {
        const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4IO<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16 *p, float (&v)[4]) { Vec<__nv_bfloat16, 4>::load(p, v); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float (&v)[4]) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 r;
        r.x = *reinterpret_cast<const uint32_t *>(&a);
        r.y = *reinterpret_cast<const uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = r;
    }
};
template <> struct Vec4IO<__half> {
    static __device__ __forceinline__ void ld(const __half *p, float (&v)[4]) { Vec<__half, 4>::load(p, v); }
    static __device__ __forceinline__ void st(__half *p, const float (&v)[4]) {
        const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
        uint2 r;
        r.x = *reinterpret_cast<const uint32_t *>(&a);
        r.y = *reinterpret_cast<const uint32_t *>(&b);
        *reinterpret_cast<uint2 *>(p) = r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exact (erf) GELU, nn.GELU() default.  erff() is ~40 instructions with branches and these kernels evaluate it per
// element, so erf is computed with Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. at fp32 rounding level of the
// GELU output; far inside the 1e-4 parity tolerance): 1 - (a1 t + .. + a5 t^5) exp(-z^2), t = 1/(1 + p z).
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = 1.f - poly * t * __expf(-z * z);   // erf(|x|/sqrt2)
    return 0.5f * x + 0.5f * fabsf(x) * e;             // 0.5 x (1 + sign(x) erf(|x|/sqrt2))
}

// w_t: depthwise weights transposed to [9][C] fp32 (tap-major), bias / ln_w / ln_b fp32 [C]
template <typename T, int J>   // J = C / 128
__global__ void __launch_bounds__(256)
dwconv3x3_ln_gelu_kernel(const T *__restrict__ x, const float *__restrict__ w_t, const float *__restrict__ bias,
                         const float *__restrict__ ln_w, const float *__restrict__ ln_b, T *__restrict__ out, int H, int W,
                         long long rows, float eps) {
    constexpr int C = J * 128;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp; pix < rows; pix += nwarps) {
        const int xw = (int)(pix % W), yh = (int)((pix / W) % H);
        const T *center = x + pix * C;
        float acc[J][4];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + 4 * (lane + 32 * j)));
            acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w;
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            if ((unsigned)(yh + dy) >= (unsigned)H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if ((unsigned)(xw + dx) >= (unsigned)W) continue;
                const T *src = center + ((long long)dy * W + dx) * C;
                const float *wt = w_t + ((dy + 1) * 3 + (dx + 1)) * C;   // Conv2d weight[c, 0, ky, kx]: cross-correlation
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float v[4];
                    Vec4IO<T>::ld(src + 4 * (lane + 32 * j), v);
                    const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wt + 4 * (lane + 32 * j)));
                    acc[j][0] = fmaf(v[0], w4.x, acc[j][0]);
                    acc[j][1] = fmaf(v[1], w4.y, acc[j][1]);
                    acc[j][2] = fmaf(v[2], w4.z, acc[j][2]);
                    acc[j][3] = fmaf(v[3], w4.w, acc[j][3]);
                }
            }
        }
        // LayerNorm over the C channels of this pixel (biased variance, eps inside the sqrt: nn.LayerNorm)
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) s += (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
        const float mean = warp_sum(s) * (1.f / C);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float d = acc[j][k] - mean;
                ss = fmaf(d, d, ss);
            }
        const float rstd = rsqrtf(warp_sum(ss) * (1.f / C) + eps);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 g4 = __ldg(reinterpret_cast<const float4 *>(ln_w + 4 * (lane + 32 * j)));
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ln_b + 4 * (lane + 32 * j)));
            float o[4];
            o[0] = gelu_erf(fmaf((acc[j][0] - mean) * rstd, g4.x, b4.x));
            o[1] = gelu_erf(fmaf((acc[j][1] - mean) * rstd, g4.y, b4.y));
            o[2] = gelu_erf(fmaf((acc[j][2] - mean) * rstd, g4.z, b4.z));
            o[3] = gelu_erf(fmaf((acc[j][3] - mean) * rstd, g4.w, b4.w));
            Vec4IO<T>::st(out + pix * C + 4 * (lane + 32 * j), o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm on channel-last activations (N, HW, C), G groups of cg = C/G channels.
//   pass 1 (gn_stats): one CTA per (n, slab of pixels): per-group partial sum / sum of squares accumulated in fp32 into
//           stats[n][g][2] with one atomicAdd pair per (CTA, group)  (stats must be zero on entry)
//   pass 2 (gn_apply): y = act((x - mean) * rstd * gamma + beta)
// The bilinear x2 upsampling that follows GN+GELU in the decoder is its own pass (upsample2x_kernel): fusing it into
// gn_apply evaluates GELU four times per output element and was measured 5x slower (compute-bound on erf).
// ---------------------------------------------------------------------------------------------------
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T *__restrict__ x, float *__restrict__ stats, int HW, int C, int G, int pix_per_cta) {
    // thread t owns channel quad (t % (C/4)) and strides over the slab's pixels
    extern __shared__ float s_part[];   // [G][2]
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    const int cg = C / G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_part[i] = 0.f;
    __syncthreads();
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    float s = 0.f, ss = 0.f;
    if (prow < pstep)
        for (int p = p0 + prow; p < p1; p += pstep) {
            float v[4];
            Vec4IO<T>::ld(x + ((long long)n * HW + p) * C + 4 * cq, v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s += v[k];
                ss = fmaf(v[k], v[k], ss);
            }
        }
    const int g = (4 * cq) / cg;   // cg is a multiple of 4 (checked on the host)
    atomicAdd(&s_part[2 * g], s);
    atomicAdd(&s_part[2 * g + 1], ss);
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&stats[(long long)n * 2 * G + i], s_part[i]);
}

// One CTA = a slab of pixels of ONE image; thread = one channel quad (fixed for the whole slab, so the group
// statistics / affine are folded into a per-thread scale+shift once) striding over the slab's pixels.
template <typename T, int ACT>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ gamma,
                const float *__restrict__ beta, T *__restrict__ y, int HW, int C, int G, float eps, int pix_per_cta) {
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int cg = C / G, g = (4 * cq) / cg;
    const float inv_cnt = 1.f / ((float)HW * cg);
    const float s = stats[((long long)n * G + g) * 2], ss = stats[((long long)n * G + g) * 2 + 1];
    const float mean = s * inv_cnt;
    const float rstd = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + eps);
    const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma + 4 * cq));
    const float4 b4 = __ldg(reinterpret_cast<const float4 *>(beta + 4 * cq));
    const float sc[4] = {rstd * g4.x, rstd * g4.y, rstd * g4.z, rstd * g4.w};
    const float sh[4] = {b4.x - mean * sc[0], b4.y - mean * sc[1], b4.z - mean * sc[2], b4.w - mean * sc[3]};
    const T *img = x + (long long)n * HW * C + 4 * cq;
    T *out = y + (long long)n * HW * C + 4 * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        float v[4], o[4];
        Vec4IO<T>::ld(img + (long long)p * C, v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t = fmaf(v[k], sc[k], sh[k]);
            o[k] = ACT == ACT_RELU ? fmaxf(t, 0.f) : ACT == ACT_GELU ? gelu_erf(t) : t;
        }
        Vec4IO<T>::st(out + (long long)p * C, o);
    }
}

// nn.UpsamplingBilinear2d(scale_factor=2) on channel-last activations: align_corners=True, src = dst * (in-1)/(out-1)
// (xyz_head.py:262-265).  One CTA = a slab of output pixels of one image, thread = one channel quad.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W, int C, int pix_per_cta) {
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int Ho = 2 * H, Wo = 2 * W;
    const float ry = (float)(H - 1) / (float)(Ho - 1), rx = (float)(W - 1) / (float)(Wo - 1);
    const T *img = x + (long long)n * H * W * C + 4 * cq;
    T *out = y + (long long)n * Ho * Wo * C + 4 * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(Ho * Wo, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        const int oh = p / Wo, ow = p - oh * Wo;
        const float fy = (float)oh * ry, fx = (float)ow * rx;
        const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        float a[4], b[4], c[4], d[4], o[4];
        Vec4IO<T>::ld(img + ((long long)y0 * W + x0) * C, a);
        Vec4IO<T>::ld(img + ((long long)y0 * W + x1) * C, b);
        Vec4IO<T>::ld(img + ((long long)y1 * W + x0) * C, c);
        Vec4IO<T>::ld(img + ((long long)y1 * W + x1) * C, d);
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // torch: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
            const float top = (1.f - lx) * a[k] + lx * b[k], bot = (1.f - lx) * c[k] + lx * d[k];
            o[k] = (1.f - ly) * top + ly * bot;
        }
        Vec4IO<T>::st(out + (long long)p * C, o);
    }
}

// MaxPool2d(kernel 3, stride 2, pad 1) on channel-last activations (network/resnet.py:106 in the stand-in backbone).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W, int C, int pix_per_cta, float floor_val) {
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const T *img = x + (long long)n * H * W * C + 4 * cq;
    T *out = y + (long long)n * Ho * Wo * C + 4 * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(Ho * Wo, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        const int oh = p / Wo, ow = p - oh * Wo;
        float m[4] = {floor_val, floor_val, floor_val, floor_val};   // 0 folds a preceding ReLU into the pool
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int ih = 2 * oh - 1 + dy;
            if ((unsigned)ih >= (unsigned)H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int iw = 2 * ow - 1 + dx;
                if ((unsigned)iw >= (unsigned)W) continue;
                float v[4];
                Vec4IO<T>::ld(img + ((long long)ih * W + iw) * C, v);
#pragma unroll
                for (int k = 0; k < 4; ++k) m[k] = fmaxf(m[k], v[k]);
            }
        }
        Vec4IO<T>::st(out + (long long)p * C, m);
    }
}

// ---------------------------------------------------------------------------------------------------
// Pose decode, one thread per RoI.  rot6 (B,6), t (B,3) = (dx, dy, z_rel) straight from fc_r / fc_t / fc_z,
// cam (B,3,3) or (1,3,3) row-major, centers (B,2), whs (B,2), ratios (B); out: rot (B,3,3) ego, trans (B,3).
// The allo->ego step is evaluated in double like the reference's numpy code (utils.py:49-60), then stored as fp32.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pose_decode_kernel(const float *__restrict__ rot6, const float *__restrict__ t, const float *__restrict__ cam, int cam_stride,
                   const float *__restrict__ centers, const float *__restrict__ whs, const float *__restrict__ ratios,
                   float *__restrict__ rot_out, float *__restrict__ trans_out, int B, int is_allo, float z_calib) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    // rot6d_to_mat_batch: x = norm(a1), z = norm(x cross a2), y = z cross x, R = [x y z] as columns (F.normalize eps 1e-12)
    const float a1[3] = {rot6[i * 6], rot6[i * 6 + 1], rot6[i * 6 + 2]}, a2[3] = {rot6[i * 6 + 3], rot6[i * 6 + 4], rot6[i * 6 + 5]};
    const float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    const float x[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float z[3] = {x[1] * a2[2] - x[2] * a2[1], x[2] * a2[0] - x[0] * a2[2], x[0] * a2[1] - x[1] * a2[0]};
    const float n3 = fmaxf(sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]), 1e-12f);
    z[0] /= n3; z[1] /= n3; z[2] /= n3;
    const float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
    float R[9] = {x[0], y[0], z[0], x[1], y[1], z[1], x[2], y[2], z[2]};

    // pose_from_predictions_test :78-119 (z_type REL; z_calib = fx/590 for wild6d, else 1)
    const float *K = cam + (long long)i * cam_stride;
    const float cx = t[i * 3] * whs[i * 2] + centers[i * 2];
    const float cy = t[i * 3 + 1] * whs[i * 2 + 1] + centers[i * 2 + 1];
    const float zz = t[i * 3 + 2] * ratios[i] * z_calib;
    const float tr[3] = {zz * (cx - K[2]) / K[0], zz * (cy - K[5]) / K[4], zz};
    trans_out[i * 3] = tr[0]; trans_out[i * 3 + 1] = tr[1]; trans_out[i * 3 + 2] = tr[2];

    if (is_allo) {   // allocentric_to_egocentric(src 'mat', dst 'mat', cam_ray (0,0,1))
        const double tx = tr[0], ty = tr[1], tz = tr[2];
        const double nt = sqrt(tx * tx + ty * ty + tz * tz);
        const double ox = tx / nt, oy = ty / nt, oz = tz / nt;
        const double angle = acos(oz);
        if (angle > 0) {
            // axis = cam_ray x obj_ray = (-oy, ox, 0); transforms3d axangle2mat normalises it
            double ax = -oy, ay = ox;
            const double na = sqrt(ax * ax + ay * ay);
            ax /= na; ay /= na;
            const double c = cos(angle), s = sin(angle), Cc = 1 - c;
            const double M[9] = {ax * ax * Cc + c, ax * ay * Cc, ay * s, ax * ay * Cc, ay * ay * Cc + c, -ax * s, -ay * s, ax * s, c};
            double E[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    E[r * 3 + cc] = M[r * 3] * (double)R[cc] + M[r * 3 + 1] * (double)R[3 + cc] + M[r * 3 + 2] * (double)R[6 + cc];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = (float)E[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) rot_out[i * 9 + k] = R[k];
}

}  // namespace gp
