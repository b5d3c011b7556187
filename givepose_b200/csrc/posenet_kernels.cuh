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

#include "gelu_fast.cuh"
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
// First MAPEncoder layer (conv_pnp_net.py:259-272): DCNv3_C's 1x1 conv maps the K = 3 NOCS channels to C, and both
// consumers of its output are linear in it, so the C-channel tensor y = conv(x) is never materialised:
//   small_k_linear   input_proj(conv(x)) = (W_ip W_c) x + (W_ip b_c + b_ip): a K-input linear map per pixel row
//                    (weights composed on the host in fp32), write-bound.
//   smallk_dwconv_ln_gelu   GELU(LN(DWConv3x3(conv(x)))) for the first `rows` pixels:
//                    dw(y)[c] = b_dw[c] + sum_{valid taps t} ( sum_k (w_dw[c,t] W_c[c,k]) x[t,k] + w_dw[c,t] b_c[c] )
//                    (zero padding applies to y, so the conv bias only enters through the taps inside the image).
//                    w_eff: [9][K+1][C] fp32, w_eff[t][k][c] = w_dw[c,t] W_c[c,k] for k < K, w_eff[t][K][c] = w_dw[c,t] b_c[c].
// ---------------------------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256)
small_k_linear_kernel(const T *__restrict__ x /*(rows,K)*/, const float *__restrict__ w /*[K][C]*/, const float *__restrict__ bias,
                      T *__restrict__ out /*(rows,C)*/, long long rows, int C) {
    constexpr int V = 16 / (int)sizeof(T);
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    float wr[K][V], b[V];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < V; ++j) wr[k][j] = __ldg(w + k * C + V * cq + j);
#pragma unroll
    for (int j = 0; j < V; ++j) b[j] = __ldg(bias + V * cq + j);
    constexpr int U = 4;   // rows in flight per thread: the kernel is a pure 16-byte store stream
    const long long stride = (long long)gridDim.x * pstep;
    for (long long r0 = (long long)blockIdx.x * pstep + prow; r0 < rows; r0 += U * stride) {
        float xv[U][K];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) xv[u][k] = r0 + u * stride < rows ? to_acc<T>(x[(r0 + u * stride) * K + k]) : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r0 + u * stride >= rows) break;
            float o[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                float a = b[j];
#pragma unroll
                for (int k = 0; k < K; ++k) a = fmaf(xv[u][k], wr[k][j], a);
                o[j] = a;
            }
            Vec<T, V>::store_stream(out + (r0 + u * stride) * C + V * cq, o);
        }
    }
}

template <typename T, int J, int K, int PX>   // J = C / 128; one warp = PX consecutive pixels of a row (W % PX == 0)
__global__ void __launch_bounds__(256)
smallk_dwconv_ln_gelu_kernel(const T *__restrict__ x /*(N,H,W,K)*/, const float *__restrict__ w_eff, const float *__restrict__ bias,
                             const float *__restrict__ ln_w, const float *__restrict__ ln_b, T *__restrict__ out, int H, int W,
                             long long rows, float eps) {
    constexpr int C = J * 128;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long pix = warp * PX; pix < rows; pix += nwarps * PX) {
        const int xw = (int)(pix % W), yh = (int)((pix / W) % H);
        float acc[PX][J][4];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + 4 * (lane + 32 * j)));
#pragma unroll
            for (int q = 0; q < PX; ++q) { acc[q][j][0] = b4.x; acc[q][j][1] = b4.y; acc[q][j][2] = b4.z; acc[q][j][3] = b4.w; }
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            if ((unsigned)(yh + dy) >= (unsigned)H) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                // the tap's K inputs of each of the PX pixels (0 outside the image; the trailing 1 carries the conv bias)
                float xv[PX][K + 1];
#pragma unroll
                for (int q = 0; q < PX; ++q) {
                    const bool ok = (unsigned)(xw + q + dx) < (unsigned)W;
                    const T *src = x + (pix + q + (long long)dy * W + dx) * K;
#pragma unroll
                    for (int k = 0; k < K; ++k) xv[q][k] = ok ? to_acc<T>(src[k]) : 0.f;
                    xv[q][K] = ok ? 1.f : 0.f;
                }
                const float *wt = w_eff + ((dy + 1) * 3 + (dx + 1)) * (K + 1) * C;
#pragma unroll
                for (int k = 0; k <= K; ++k)
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wt + k * C + 4 * (lane + 32 * j)));
#pragma unroll
                        for (int q = 0; q < PX; ++q) {
                            acc[q][j][0] = fmaf(xv[q][k], w4.x, acc[q][j][0]);
                            acc[q][j][1] = fmaf(xv[q][k], w4.y, acc[q][j][1]);
                            acc[q][j][2] = fmaf(xv[q][k], w4.z, acc[q][j][2]);
                            acc[q][j][3] = fmaf(xv[q][k], w4.w, acc[q][j][3]);
                        }
                    }
            }
        }
#pragma unroll
        for (int q = 0; q < PX; ++q) {
            if (pix + q >= rows) break;
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) s += (acc[q][j][0] + acc[q][j][1]) + (acc[q][j][2] + acc[q][j][3]);
            const float mean = warp_sum(s) * (1.f / C);
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d = acc[q][j][k] - mean;
                    ss = fmaf(d, d, ss);
                }
            const float rstd = rsqrtf(warp_sum(ss) * (1.f / C) + eps);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const float4 g4 = __ldg(reinterpret_cast<const float4 *>(ln_w + 4 * (lane + 32 * j)));
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(ln_b + 4 * (lane + 32 * j)));
                float o[4];
                o[0] = gelu_erf(fmaf((acc[q][j][0] - mean) * rstd, g4.x, b4.x));
                o[1] = gelu_erf(fmaf((acc[q][j][1] - mean) * rstd, g4.y, b4.y));
                o[2] = gelu_erf(fmaf((acc[q][j][2] - mean) * rstd, g4.z, b4.z));
                o[3] = gelu_erf(fmaf((acc[q][j][3] - mean) * rstd, g4.w, b4.w));
                Vec4IO<T>::st(out + (pix + q) * C + 4 * (lane + 32 * j), o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// First MAPEncoder layer, whole DCNv3 module in one kernel (conv_pnp_net.py:259-272, modules/dcnv3.py:318-356).
// The module is  output_proj( core( input_proj(conv1x1(x)), offset, softmax(mask) ) )  with x the K = 3 channel NOCS map.
// conv1x1, input_proj and output_proj are linear and the core is linear in its input, so for output pixel q
//   core[q, g, c] = sum_j Wp[g*gc+c, j] * S[q,g,j] + bp[g*gc+c] * S0[q,g]
//   S[q,g,j] = sum_p m_p sum_{corners k inside} w_k x[corner_k, j]      S0[q,g] = sum_p m_p sum_{corners k inside} w_k
// (Wp, bp = input_proj o conv1x1; a corner outside the image contributes neither pixel value nor bias, cuh:55-75), hence
//   out[q, :] = W2^T (S[q,0,0..2], S0[q,0], ..., S[q,G-1,0..2], S0[q,G-1]) + b_out,  W2 [G*(K+1)][C] composed on the host in fp64.
// The C-channel input_proj tensor (2.1 GB per 1024 RoIs), the core output and the output_proj GEMM disappear; the sampling
// arithmetic per (pixel, group, point) is locate() -- the same function the tiled sampler and the index hook use.
// One warp = PX output pixels per iteration: phase 1 lane-linear over the PX*G*P (pixel, group, point) triples (partial sums
// in shared memory), phase 2 the G*(K+1) sums per pixel, phase 3 the [G*(K+1)] x C map, lane = C/32 consecutive channels.
// ---------------------------------------------------------------------------------------------------
struct SmallKFusedParams {
    int H, W, Ho, Wo, sh, sw, dh, dw, base_h, base_w, half_h, half_w;
    float scale;
    long long n_pix;   // N*Ho*Wo
};

template <typename T, int PX>
__global__ void __launch_bounds__(256, 3)
dcnv3_smallk_fused_kernel(const T *__restrict__ x /*(N,H,W,3)*/, const T *__restrict__ off, const T *__restrict__ msk /*logits*/,
                          const float *__restrict__ w2 /*[16][256]*/, const float *__restrict__ bias /*[256]*/, T *__restrict__ out,
                          const __grid_constant__ SmallKFusedParams p) {
    constexpr int G = 4, P = 9, K = 3, NS = G * (K + 1), C = 256, CPL = C / 32, NT = PX * G * P;
    __shared__ __align__(16) float s_w2[NS * C];
    __shared__ float s_part[8][PX * G * P][K + 2];   // per warp: (sum e w v_j, sum e w, e) per triple
    __shared__ float s_S[8][PX][NS];
    for (int i = threadIdx.x; i < NS * C; i += blockDim.x) s_w2[i] = w2[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float bs[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) bs[k] = __ldg(bias + CPL * lane + k);
    const long long n_groups = (p.n_pix + PX - 1) / PX;
    const long long HoWo = (long long)p.Ho * p.Wo;
    for (long long grp = (long long)blockIdx.x * 8 + warp; grp < n_groups; grp += (long long)gridDim.x * 8) {
        const long long q0 = grp * PX;
        // ---- phase 1: one (pixel, group, point) triple per lane per pass
        for (int t = lane; t < NT; t += 32) {
            const int pix = t / (G * P), rem = t - pix * (G * P), g = rem / P, pt = rem - g * P;
            const long long q = q0 + pix;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, e = 0.f;
            if (q < p.n_pix) {
                const long long b = q / HoWo;
                const int r = (int)(q - b * HoWo), oh = r / p.Wo, ow = r - oh * p.Wo;
                const long long row = (q * G + g) * P;   // flat (q*G+g)*P addressing, cuh:243-244
                const T *mrow = msk + row;
                float mx = to_acc<T>(__ldg(mrow));
#pragma unroll
                for (int i = 1; i < P; ++i) mx = fmaxf(mx, to_acc<T>(__ldg(mrow + i)));
                e = expf(to_acc<T>(__ldg(mrow + pt)) - mx);   // softmax numerator; the denominator is summed in phase 2
                float ox, oy;
                load_pair<T>(off + 2 * (row + pt), ox, oy);
                const float p0_h_ = origin<float>(p.base_h + oh * p.sh, p.half_h, p.scale);
                const float p0_w_ = origin<float>(p.base_w + ow * p.sw, p.half_w, p.scale);
                const int i = pt / 3, j = pt - 3 * i;   // p = i*kh + j, kernel WIDTH index slow (cuh:257-258)
                Point<float> sp;
                locate<float>(sp, p0_h_, p0_w_, j * p.dh, i * p.dw, ox, oy, p.scale, p.H, p.W);
                if (sp.flags & F_IN) {
                    const T *img = x + (b * p.H * p.W) * K;
                    const float w1 = sp.hh * sp.hw, w2_ = sp.hh * sp.lw, w3 = sp.lh * sp.hw, w4 = sp.lh * sp.lw;
                    auto corner = [&](unsigned flag, int yy, int xx, float w) {
                        if (sp.flags & flag) {
                            const T *px = img + ((long long)yy * p.W + xx) * K;
                            a0 = fmaf(w, to_acc<T>(__ldg(px)), a0);
                            a1 = fmaf(w, to_acc<T>(__ldg(px + 1)), a1);
                            a2 = fmaf(w, to_acc<T>(__ldg(px + 2)), a2);
                            a3 += w;
                        }
                    };
                    corner(F_C1, sp.h_low, sp.w_low, w1);
                    corner(F_C2, sp.h_low, sp.w_low + 1, w2_);
                    corner(F_C3, sp.h_low + 1, sp.w_low, w3);
                    corner(F_C4, sp.h_low + 1, sp.w_low + 1, w4);
                }
            }
            float *sp_ = s_part[warp][t];
            sp_[0] = e * a0; sp_[1] = e * a1; sp_[2] = e * a2; sp_[3] = e * a3; sp_[4] = e;
        }
        __syncwarp();
        // ---- phase 2: the NS sums of each pixel, normalised by the softmax denominator of their group
        for (int t = lane; t < PX * NS; t += 32) {
            const int pix = t / NS, s = t - pix * NS, g = s >> 2, jj = s & 3;
            const float(*pp)[K + 2] = &s_part[warp][(pix * G + g) * P];
            float num = 0.f, den = 0.f;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                num += pp[k][jj];
                den += pp[k][4];
            }
            s_S[warp][pix][s] = den > 0.f ? num / den : 0.f;   // den == 0 only for pixels past n_pix
        }
        __syncwarp();
        // ---- phase 3: out[q, c] = b[c] + sum_s S[q,s] W2[s][c] for the lane's CPL channels of the PX pixels
        float acc[PX][CPL];
#pragma unroll
        for (int px = 0; px < PX; ++px)
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[px][k] = bs[k];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            float w[CPL];
#pragma unroll
            for (int k4 = 0; k4 < CPL; k4 += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&s_w2[s * C + CPL * lane + k4]);
                w[k4] = v.x; w[k4 + 1] = v.y; w[k4 + 2] = v.z; w[k4 + 3] = v.w;
            }
#pragma unroll
            for (int px = 0; px < PX; ++px) {
                const float sv = s_S[warp][px][s];
#pragma unroll
                for (int k = 0; k < CPL; ++k) acc[px][k] = fmaf(sv, w[k], acc[px][k]);
            }
        }
#pragma unroll
        for (int px = 0; px < PX; ++px) {
            const long long q = q0 + px;
            if (q < p.n_pix) {
                constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
                for (int k = 0; k < CPL; k += V)
                    Vec<T, V>::store_stream(out + q * C + CPL * lane + k, *reinterpret_cast<float(*)[V]>(&acc[px][k]));
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm on channel-last activations (N, HW, C), G groups of cg = C/G channels.
//   pass 1 (gn_stats): one CTA per (n, slab of pixels): per-group partial sum / sum of squares in fp32, written to
//           partial[n][slab][g][2]; gn_finalize turns them into (mean, rstd) per (n, g).  No atomics anywhere: sums are taken
//           in a fixed order, so the result is bit-reproducible.
//   pass 2 (gn_apply): y = act((x - mean) * rstd * gamma + beta)
// The bilinear x2 upsampling that follows GN+GELU in the decoder is its own pass (upsample2x_kernel): fusing it into
// gn_apply evaluates GELU four times per output element and was measured 5x slower (compute-bound on erf).
// ---------------------------------------------------------------------------------------------------
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T *__restrict__ x, float *__restrict__ partial /*[N][slabs][G][2]*/, int HW, int C, int G, int pix_per_cta) {
    // thread t owns channel quad (t % (C/4)) and strides over the slab's pixels; everything is summed in a fixed order
    // (no atomics): the forward is bit-reproducible run to run
    extern __shared__ float s_thr[];   // [blockDim][2]
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    const int cg = C / G;
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
    s_thr[2 * threadIdx.x] = s;
    s_thr[2 * threadIdx.x + 1] = ss;
    __syncthreads();
    // thread i < 2G: statistic (i & 1) of group (i >> 1) = its cg/4 quads x pstep pixel rows, in index order
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
        const int g = i >> 1, qpg = cg / 4;   // cg is a multiple of 4 (checked on the host)
        float a = 0.f;
        for (int r = 0; r < pstep; ++r)
            for (int j = 0; j < qpg; ++j) a += s_thr[2 * (r * q + g * qpg + j) + (i & 1)];
        partial[(((long long)n * gridDim.x + blockIdx.x) * G + g) * 2 + (i & 1)] = a;
    }
}

// (mean, rstd) per (n, group) from the slab partials, summed in slab order
__global__ void __launch_bounds__(128)
gn_finalize_kernel(const float *__restrict__ partial, float *__restrict__ stats /*[N][G][2]*/, int NG, int G, int slabs, float inv_cnt,
                   float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NG) return;
    const int n = i / G, g = i - n * G;
    float s = 0.f, ss = 0.f;
    for (int k = 0; k < slabs; ++k) {
        const float *pp = partial + (((long long)n * slabs + k) * G + g) * 2;
        s += pp[0];
        ss += pp[1];
    }
    const float mean = s * inv_cnt;
    stats[2 * i] = mean;
    stats[2 * i + 1] = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + eps);
}

// GELU for 16-bit storage: 0.5 x (1 + tanh(x (a + b x^2 + c x^4))) with (a, b, c) fitted to the exact erf GELU
// (max abs deviation 2.5e-5 on [-8, 8]; the argument is clamped to +-6 where tanh has saturated) and the hardware
// tanh.approx.f32 (rel. error 2^-11): total error < 2.5e-4 |x|, an order of magnitude below the bf16 rounding of the
// stored result.  9 instructions / 1 MUFU instead of ~18 / 2: gn_apply is otherwise bound by the GELU arithmetic, not
// by HBM.  The fp32 parity mode keeps gelu_erf.

template <typename T, int ACT> __device__ __forceinline__ float gn_act(float t) {
    if (ACT == ACT_RELU) return fmaxf(t, 0.f);
    if (ACT == ACT_GELU) return sizeof(T) == 2 ? gelu_fast16(t) : gelu_erf(t);
    return t;
}

// per-thread folded GroupNorm affine of V consecutive channels starting at c0: y = x * sc + sh
template <int V>
__device__ __forceinline__ void gn_fold(const float *__restrict__ stats, const float *__restrict__ gamma,
                                        const float *__restrict__ beta, int n, int c0, int C, int G, int HW, float eps,
                                        float (&sc)[V], float (&sh)[V]) {
    const int cg = C / G;
    (void)HW; (void)eps;   // folded into stats = (mean, rstd) by gn_finalize_kernel
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = c0 + k, g = c / cg;
        const float mean = stats[((long long)n * G + g) * 2], rstd = stats[((long long)n * G + g) * 2 + 1];
        sc[k] = rstd * __ldg(gamma + c);
        sh[k] = __ldg(beta + c) - mean * sc[k];
    }
}

// One CTA = a slab of pixels of ONE image; thread = V = 16 bytes of channels (fixed for the whole slab, so the group
// statistics / affine are folded into per-thread scale+shift once) striding over the slab's pixels, two pixels in flight.
template <typename T, int ACT>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const T *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ gamma,
                const float *__restrict__ beta, T *__restrict__ y, int HW, int C, int G, float eps, int pix_per_cta) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.y;
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    float sc[V], sh[V];
    gn_fold<V>(stats, gamma, beta, n, V * cq, C, G, HW, eps, sc, sh);
    const T *img = x + (long long)n * HW * C + V * cq;
    T *out = y + (long long)n * HW * C + V * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += 2 * pstep) {
        const bool two = p + pstep < p1;
        float v0[V], v1[V];
        Vec<T, V>::load(img + (long long)p * C, v0);
        if (two) Vec<T, V>::load(img + (long long)(p + pstep) * C, v1);
#pragma unroll
        for (int k = 0; k < V; ++k) v0[k] = gn_act<T, ACT>(fmaf(v0[k], sc[k], sh[k]));
        Vec<T, V>::store_stream(out + (long long)p * C, v0);
        if (two) {
#pragma unroll
            for (int k = 0; k < V; ++k) v1[k] = gn_act<T, ACT>(fmaf(v1[k], sc[k], sh[k]));
            Vec<T, V>::store_stream(out + (long long)(p + pstep) * C, v1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward of y = act(GroupNorm(x)) for the training step (replaces torch's native_group_norm_backward + gelu_backward /
// threshold_backward + the fp32 <-> bf16 copies autocast wraps around them).  With z = xhat*gamma + beta,
// xhat = (x - mean)*rstd, dz = dy * act'(z) (z is recomputed, nothing but x and the (mean, rstd) pairs is saved):
//   dgamma[c] = sum_{n,p} dz*xhat      dbeta[c] = sum_{n,p} dz
//   A[n,g] = sum_{p, c in g} dz*gamma  B[n,g] = sum_{p, c in g} dz*gamma*xhat
//   dx = rstd * (dz*gamma - (A + xhat*B) / (HW*cg))
//   pass 1 (gn_bwd_stats):   per (n, slab, c): sum dz, sum dz*xhat  -> partial[n][slab][c][2]  (fixed order, no atomics)
//   finalize (gn_bwd_group): per (n, g): A/cnt, B/cnt -> gstat[n][g][2];  (gn_bwd_param): one warp per channel -> dgamma, dbeta
//   pass 2 (gn_bwd_apply):   dx
// ---------------------------------------------------------------------------------------------------
template <typename T, int ACT> __device__ __forceinline__ float gn_act_grad(float z) {
    if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
    if (ACT == ACT_GELU && sizeof(T) == 2) {
        // 16-bit storage: the derivative of exactly what the forward evaluates (gelu_fast16): g = 0.5 z (1 + tanh u),
        // u = zc (a + b zc^2 + c zc^4), zc = clamp(z, +-6):  g' = 0.5 (1 + t) + 0.5 z (1 - t^2) u'(zc)   (u' = 0 where clamped)
        const float zc = fminf(fmaxf(z, -6.f), 6.f);
        const float z2 = zc * zc;
        const float u = zc * fmaf(z2, fmaf(z2, -3.51523083e-4f, 3.70056758e-2f), 7.97507859e-1f);
        const float du = fabsf(z) < 6.f ? fmaf(z2, fmaf(z2, 5.f * -3.51523083e-4f, 3.f * 3.70056758e-2f), 7.97507859e-1f) : 0.f;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
        return fmaf(0.5f * z * du, fmaf(-t, t, 1.f), fmaf(0.5f, t, 0.5f));
    }
    if (ACT == ACT_GELU) {   // d/dz [z Phi(z)] = Phi(z) + z phi(z); erf as in gelu_erf (A&S 7.1.26), exp(-z^2/2) shared
        const float a = fabsf(z) * 0.70710678118654752440f;
        const float t = __frcp_rn(fmaf(0.3275911f, a, 1.f));
        float poly = fmaf(1.061405429f, t, -1.453152027f);
        poly = fmaf(poly, t, 1.421413741f);
        poly = fmaf(poly, t, -0.284496736f);
        poly = fmaf(poly, t, 0.254829592f);
        const float E = __expf(-0.5f * z * z);
        const float erf_abs = 1.f - poly * t * E;
        const float Phi = 0.5f + copysignf(0.5f * erf_abs, z);
        return fmaf(z * 0.39894228040143267794f, E, Phi);
    }
    return 1.f;
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
gn_bwd_stats_kernel(const T *__restrict__ x, const T *__restrict__ dy, const float *__restrict__ stats, const float *__restrict__ gamma,
                    const float *__restrict__ beta, float *__restrict__ partial /*[N][slabs][C][2]*/, int HW, int C, int G,
                    int pix_per_cta) {
    extern __shared__ float s_thr[];   // [blockDim][8]
    const int n = blockIdx.y;
    const int q = C / 4, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    const int cg = C / G;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (prow < pstep) {
        float sc[4], sh[4], mean[4], rstd[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 4 * cq + k, g = c / cg;
            mean[k] = stats[((long long)n * G + g) * 2];
            rstd[k] = stats[((long long)n * G + g) * 2 + 1];
            sc[k] = rstd[k] * __ldg(gamma + c);
            sh[k] = __ldg(beta + c) - mean[k] * sc[k];
        }
        for (int p = p0 + prow; p < p1; p += 2 * pstep) {   // two pixels in flight; sums stay in pixel order
            const bool two = p + pstep < p1;
            float v[4], d[4], v2[4], d2[4];
            const long long o = ((long long)n * HW + p) * C + 4 * cq;
            Vec4IO<T>::ld(x + o, v);
            Vec4IO<T>::ld(dy + o, d);
            if (two) {
                Vec4IO<T>::ld(x + o + (long long)pstep * C, v2);
                Vec4IO<T>::ld(dy + o + (long long)pstep * C, d2);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float dz = d[k] * gn_act_grad<T, ACT>(fmaf(v[k], sc[k], sh[k]));
                s1[k] += dz;
                s2[k] = fmaf(dz, (v[k] - mean[k]) * rstd[k], s2[k]);
            }
            if (two) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float dz = d2[k] * gn_act_grad<T, ACT>(fmaf(v2[k], sc[k], sh[k]));
                    s1[k] += dz;
                    s2[k] = fmaf(dz, (v2[k] - mean[k]) * rstd[k], s2[k]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s_thr[8 * threadIdx.x + 2 * k] = s1[k];
        s_thr[8 * threadIdx.x + 2 * k + 1] = s2[k];
    }
    __syncthreads();
    // entry i = (channel c, which): the pstep pixel rows of quad c/4, in row order
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        const int c = i >> 1;
        float a = 0.f;
        for (int r = 0; r < pstep; ++r) a += s_thr[8 * (r * q + (c >> 2)) + 2 * (c & 3) + (i & 1)];
        partial[(((long long)n * gridDim.x + blockIdx.x) * C + c) * 2 + (i & 1)] = a;
    }
}

__global__ void __launch_bounds__(128)
gn_bwd_group_kernel(const float *__restrict__ partial, const float *__restrict__ gamma, float *__restrict__ gstat /*[N][G][2]*/,
                    int NG, int G, int C, int slabs, float inv_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NG) return;
    const int n = i / G, g = i - n * G, cg = C / G;
    float A = 0.f, B = 0.f;
    for (int k = 0; k < slabs; ++k)
        for (int j = 0; j < cg; ++j) {
            const int c = g * cg + j;
            const float *pp = partial + (((long long)n * slabs + k) * C + c) * 2;
            const float gm = __ldg(gamma + c);
            A = fmaf(gm, pp[0], A);
            B = fmaf(gm, pp[1], B);
        }
    gstat[2 * i] = A * inv_cnt;
    gstat[2 * i + 1] = B * inv_cnt;
}

// one warp per channel: lanes stride over the (n, slab) partials, then a fixed-order butterfly
__global__ void __launch_bounds__(256)
gn_bwd_param_kernel(const float *__restrict__ partial, float *__restrict__ dgamma, float *__restrict__ dbeta, int C, int n_slabs_total) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    float b = 0.f, g = 0.f;
    for (int k = lane; k < n_slabs_total; k += 32) {
        const float2 v = *reinterpret_cast<const float2 *>(partial + ((long long)k * C + c) * 2);
        b += v.x;
        g += v.y;
    }
    b = warp_sum(b);
    g = warp_sum(g);
    if (lane == 0) {
        dbeta[c] = b;
        dgamma[c] = g;
    }
}

template <typename T, int ACT>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const T *__restrict__ x, const T *__restrict__ dy, const float *__restrict__ stats, const float *__restrict__ gstat,
                    const float *__restrict__ gamma, const float *__restrict__ beta, T *__restrict__ dx, int HW, int C, int G,
                    int pix_per_cta) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.y;
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int cg = C / G;
    // dx = dz * (rstd*gamma) - rstd*(A' + xhat*B'),  xhat = x*rstd - mean*rstd  ->  dx = dz*sc - (x*c1 + c0)
    float sc[V], sh[V], c0[V], c1[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = V * cq + k, g = c / cg;
        const float mean = stats[((long long)n * G + g) * 2], rstd = stats[((long long)n * G + g) * 2 + 1];
        const float A = gstat[((long long)n * G + g) * 2], B = gstat[((long long)n * G + g) * 2 + 1];
        sc[k] = rstd * __ldg(gamma + c);
        sh[k] = __ldg(beta + c) - mean * sc[k];
        c1[k] = rstd * rstd * B;
        c0[k] = rstd * (A - mean * rstd * B);
    }
    const long long base = (long long)n * HW * C + V * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        float v[V], d[V];
        Vec<T, V>::load(x + base + (long long)p * C, v);
        Vec<T, V>::load_stream(dy + base + (long long)p * C, d);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const float dz = d[k] * gn_act_grad<T, ACT>(fmaf(v[k], sc[k], sh[k]));
            d[k] = fmaf(dz, sc[k], -fmaf(v[k], c1[k], c0[k]));
        }
        Vec<T, V>::store_stream(dx + base + (long long)p * C, d);
    }
}

// GroupNorm -> act -> 1x1 convolution to OC (<= 4) channels + bias, for the decoder's out_layer (xyz_head.py:349-366:
// the last ConvModule's GN + GELU followed by Conv1x1 256 -> 3): the normalised 256-channel activation is never written.
// One warp per pixel, lane = C/32 consecutive channels; the OC dot products are reduced with warp shuffles.
template <typename T, int ACT, int CPL /*channels per lane*/, int OC>
__global__ void __launch_bounds__(256)
gn_act_conv1x1_kernel(const T *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ gamma,
                      const float *__restrict__ beta, const float *__restrict__ w /*[OC][C]*/, const float *__restrict__ bias,
                      T *__restrict__ y /*(N*HW, OC)*/, int HW, int G, float eps, int pix_per_cta) {
    constexpr int C = CPL * 32, V = 16 / (int)sizeof(T), NV = CPL / V;
    static_assert(CPL % V == 0, "a lane owns whole 16-byte vectors");
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float sc[CPL], sh[CPL], wr[OC][CPL];
    gn_fold<CPL>(stats, gamma, beta, n, CPL * lane, C, G, HW, eps, sc, sh);
#pragma unroll
    for (int o = 0; o < OC; ++o)
#pragma unroll
        for (int k = 0; k < CPL; ++k) wr[o][k] = __ldg(w + o * C + CPL * lane + k);
    const float b = lane < OC ? __ldg(bias + lane) : 0.f;
    const T *img = x + (long long)n * HW * C + CPL * lane;
    T *out = y + (long long)n * HW * OC;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    for (int p = p0 + warp; p < p1; p += nwarp) {
        float v[CPL];
#pragma unroll
        for (int j = 0; j < NV; ++j) Vec<T, V>::load(img + (long long)p * C + j * V, *reinterpret_cast<float(*)[V]>(&v[j * V]));
        float acc[OC];
#pragma unroll
        for (int o = 0; o < OC; ++o) acc[o] = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
            const float a = gn_act<T, ACT>(fmaf(v[k], sc[k], sh[k]));
#pragma unroll
            for (int o = 0; o < OC; ++o) acc[o] = fmaf(a, wr[o][k], acc[o]);
        }
        float mine = 0.f;
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            const float r = warp_sum(acc[o]);
            if (lane == o) mine = r;
        }
        if (lane < OC) out[(long long)p * OC + lane] = from_acc<T, float>(mine + b);
    }
}

// nn.UpsamplingBilinear2d(scale_factor=2) on channel-last activations: align_corners=True, src = dst * (in-1)/(out-1)
// (xyz_head.py:262-265).  One CTA = a slab of output pixels of one image, thread = 16 bytes of channels.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W, int C, int pix_per_cta) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.y;
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int Ho = 2 * H, Wo = 2 * W;
    const float ry = (float)(H - 1) / (float)(Ho - 1), rx = (float)(W - 1) / (float)(Wo - 1);
    const T *img = x + (long long)n * H * W * C + V * cq;
    T *out = y + (long long)n * Ho * Wo * C + V * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(Ho * Wo, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        const int oh = p / Wo, ow = p - oh * Wo;
        const float fy = (float)oh * ry, fx = (float)ow * rx;
        const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
        const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
        const float ly = fy - (float)y0, lx = fx - (float)x0;
        float a[V], b[V], c[V], d[V], o[V];
        Vec<T, V>::load(img + ((long long)y0 * W + x0) * C, a);
        Vec<T, V>::load(img + ((long long)y0 * W + x1) * C, b);
        Vec<T, V>::load(img + ((long long)y1 * W + x0) * C, c);
        Vec<T, V>::load(img + ((long long)y1 * W + x1) * C, d);
#pragma unroll
        for (int k = 0; k < V; ++k) {   // torch: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
            const float top = (1.f - lx) * a[k] + lx * b[k], bot = (1.f - lx) * c[k] + lx * d[k];
            o[k] = (1.f - ly) * top + ly * bot;
        }
        Vec<T, V>::store_stream(out + (long long)p * C, o);
    }
}

// Column-strip variant of upsample2x_kernel for the decoder shapes (q = C/V threads per pixel divides 256, Wo a multiple of
// 256/q): a thread owns ONE output column x V channels and walks down a slab of output rows, so everything that depends on the
// column (source columns, lx, addresses) is computed once and the per-row work is 4 loads + the blend + 1 store.  The generic
// kernel spends 175 instructions per output vector (ncu: issue slots 73 % busy, DRAM 34 %), most of them index arithmetic.
// fp32 storage keeps torch's evaluation order (parity mode); 16-bit storage uses the four folded weights (the result is rounded
// to 16 bits anyway).
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_strip_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W, int C, int rows_per_cta) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.z;
    const int q = C / V, cq = threadIdx.x % q, col = threadIdx.x / q, cols = blockDim.x / q;
    const int Ho = 2 * H, Wo = 2 * W;
    const int ow = blockIdx.x * cols + col;
    const float ry = (float)(H - 1) / (float)(Ho - 1), rx = (float)(W - 1) / (float)(Wo - 1);
    const float fx = (float)ow * rx;
    const int x0 = min((int)fx, W - 1), x1 = min(x0 + 1, W - 1);
    const float lx = fx - (float)x0;
    const T *c0 = x + (long long)n * H * W * C + (long long)x0 * C + V * cq;
    const T *c1 = x + (long long)n * H * W * C + (long long)x1 * C + V * cq;
    T *out = y + (long long)n * Ho * Wo * C + (long long)ow * C + V * cq;
    const int oh0 = blockIdx.y * rows_per_cta, oh1 = min(Ho, oh0 + rows_per_cta);
    const long long rs = (long long)W * C, ors = (long long)Wo * C;
    // the two source rows stay in registers while the output row walks down: the source row index advances every second
    // output row (ry ~ 1/2), and then only ONE new row is loaded (the old lower row becomes the upper one)
    float a[V], b[V], c[V], d[V];
    int cy0 = -2, cy1 = -2;
    for (int oh = oh0; oh < oh1; ++oh) {
        const float fy = (float)oh * ry;
        const int y0 = min((int)fy, H - 1), y1 = min(y0 + 1, H - 1);
        const float ly = fy - (float)y0;
        if (y0 != cy0) {
            if (y0 == cy1) {
#pragma unroll
                for (int k = 0; k < V; ++k) { a[k] = c[k]; b[k] = d[k]; }
            } else {
                Vec<T, V>::load(c0 + y0 * rs, a);
                Vec<T, V>::load(c1 + y0 * rs, b);
            }
            cy0 = y0;
        }
        if (y1 != cy1) {
            if (y1 == y0) {
#pragma unroll
                for (int k = 0; k < V; ++k) { c[k] = a[k]; d[k] = b[k]; }
            } else {
                Vec<T, V>::load(c0 + y1 * rs, c);
                Vec<T, V>::load(c1 + y1 * rs, d);
            }
            cy1 = y1;
        }
        float o[V];
        if (sizeof(T) == 4) {
#pragma unroll
            for (int k = 0; k < V; ++k) {   // torch: h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
                const float top = (1.f - lx) * a[k] + lx * b[k], bot = (1.f - lx) * c[k] + lx * d[k];
                o[k] = (1.f - ly) * top + ly * bot;
            }
        } else {
            const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
            for (int k = 0; k < V; ++k) o[k] = fmaf(w11, d[k], fmaf(w10, c[k], fmaf(w01, b[k], w00 * a[k])));
        }
        Vec<T, V>::store_stream(out + oh * ors, o);
    }
}

// Backward of upsample2x_kernel as a GATHER (deterministic, no atomics): input pixel (iy, ix) collects
// wy(oh, iy) * wx(ow, ix) * dy[oh, ow] over the <= 6 x 6 output pixels whose 2x2 source footprint contains it; the weights
// are recomputed with exactly the forward's expressions, so forward and backward are transposes of each other bit for bit.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const T *__restrict__ dy, T *__restrict__ dx, int H, int W, int C, int pix_per_cta) {
    constexpr int V = 16 / (int)sizeof(T), R = 6;
    const int n = blockIdx.y;
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int Ho = 2 * H, Wo = 2 * W;
    const float ry = (float)(H - 1) / (float)(Ho - 1), rx = (float)(W - 1) / (float)(Wo - 1);
    const T *g = dy + (long long)n * Ho * Wo * C + V * cq;
    T *out = dx + (long long)n * H * W * C + V * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(H * W, p0 + pix_per_cta);
    auto weight = [](int o, int i, float r, int n_in) {   // contribution of output index o to input index i (forward's y0/y1/ly)
        const float f = (float)o * r;
        const int i0 = min((int)f, n_in - 1), i1 = min(i0 + 1, n_in - 1);
        const float l = f - (float)i0;
        return (i0 == i ? 1.f - l : 0.f) + (i1 == i ? l : 0.f);
    };
    for (int p = p0 + prow; p < p1; p += pstep) {
        const int iy = p / W, ix = p - iy * W;
        // outputs with source coordinate in (i-1, i+1): o in ((i-1)/r, (i+1)/r); r ~ 1/2 -> at most 5, R = 6 with slack
        const int oy0 = max(0, (int)floorf((float)(iy - 1) / fmaxf(ry, 1e-6f))), ox0 = max(0, (int)floorf((float)(ix - 1) / fmaxf(rx, 1e-6f)));
        float wy[R], wx[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
            wy[t] = oy0 + t < Ho ? weight(oy0 + t, iy, ry, H) : 0.f;
            wx[t] = ox0 + t < Wo ? weight(ox0 + t, ix, rx, W) : 0.f;
        }
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
        for (int a = 0; a < R; ++a) {
            if (wy[a] == 0.f) continue;
#pragma unroll
            for (int b = 0; b < R; ++b) {
                if (wx[b] == 0.f) continue;
                float v[V];
                Vec<T, V>::load(g + ((long long)(oy0 + a) * Wo + ox0 + b) * C, v);
                const float w = wy[a] * wx[b];
#pragma unroll
                for (int k = 0; k < V; ++k) acc[k] = fmaf(w, v[k], acc[k]);
            }
        }
        Vec<T, V>::store_stream(out + (long long)p * C, acc);
    }
}

// Input packing for the stand-in backbone's 7x7 / stride-2 / pad-3 stem (network/resnet.py:104): cuDNN has no tensor-core
// kernel worth the name for 3 input channels (10.6 ms per 1024 RoIs on B200, 16 % of the whole forward).  The same
// convolution is a 4x4 / stride-1 convolution over the 2x2 space-to-depth image (kernel padded 7 -> 8 with a zero tap in
// front): 12 -> 16 channels, K = 256 -- an ordinary implicit GEMM.  This kernel reads the fp32 NCHW RoI crops once and
// writes the packed operand: out[n, 2 + y2, 2 + x2, c*4 + ry*2 + rx] = img[n, c, 2*y2 + ry, 2*x2 + rx], zero elsewhere,
// shape (N, H/2 + 3, W/2 + 3, 16) channel-last in the compute dtype (replaces the cast + layout-change copies).
template <typename T>
__global__ void __launch_bounds__(256)
stem_s2d_pack_kernel(const float *__restrict__ img, T *__restrict__ out, int N, int H, int W) {
    const int H2 = H / 2, W2 = W / 2, Hp = H2 + 3, Wp = W2 + 3;
    const long long total = (long long)N * Hp * Wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % Wp), Y = (int)((i / Wp) % Hp);
        const long long n = i / ((long long)Wp * Hp);
        const int y2 = Y - 2, x2 = X - 2;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0.f;
        if ((unsigned)y2 < (unsigned)H2 && (unsigned)x2 < (unsigned)W2) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int ry = 0; ry < 2; ++ry) {
                    const float2 r = __ldg(reinterpret_cast<const float2 *>(img + ((n * 3 + c) * H + 2 * y2 + ry) * W + 2 * x2));
                    v[c * 4 + ry * 2] = r.x;
                    v[c * 4 + ry * 2 + 1] = r.y;
                }
        }
        T *o = out + i * 16;
        constexpr int V = 16 / (int)sizeof(T);
#pragma unroll
        for (int j = 0; j < 16 / V; ++j) Vec<T, V>::store_stream(o + j * V, *reinterpret_cast<float(*)[V]>(&v[j * V]));
    }
}

// y = relu(y + bias[c] + residual), in place, channel-last (rows, C): the tail of a residual block when the convolution
// runs without cuDNN's fused add+ReLU epilogue (whose kernels for 64 / 128 channels run at a third of the plain convolution's
// rate on B200: 0.68 ms against 0.25 ms per 1024 RoIs at 64x64x64).  One pass, 16-byte vectors.
template <typename T>
__global__ void __launch_bounds__(256)
bias_add_relu_kernel(T *__restrict__ y, const T *__restrict__ res, const float *__restrict__ bias, long long n_vec, int C) {
    constexpr int V = 16 / (int)sizeof(T);
    const int q = C / V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % q) * V;
        float a[V], r[V];
        Vec<T, V>::load_stream(y + i * V, a);
        Vec<T, V>::load_stream(res + i * V, r);
#pragma unroll
        for (int k = 0; k < V; ++k) a[k] = fmaxf(a[k] + __ldg(bias + c0 + k) + r[k], 0.f);
        Vec<T, V>::store_stream(y + i * V, a);
    }
}

// MaxPool2d(kernel 3, stride 2, pad 1) on channel-last activations (network/resnet.py:106 in the stand-in backbone).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const T *__restrict__ x, T *__restrict__ y, int H, int W, int C, int pix_per_cta, float floor_val) {
    constexpr int V = 16 / (int)sizeof(T);
    const int n = blockIdx.y;
    const int q = C / V, cq = threadIdx.x % q, prow = threadIdx.x / q, pstep = blockDim.x / q;
    if (prow >= pstep) return;
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const T *img = x + (long long)n * H * W * C + V * cq;
    T *out = y + (long long)n * Ho * Wo * C + V * cq;
    const int p0 = blockIdx.x * pix_per_cta, p1 = min(Ho * Wo, p0 + pix_per_cta);
    for (int p = p0 + prow; p < p1; p += pstep) {
        const int oh = p / Wo, ow = p - oh * Wo;
        float m[V];
#pragma unroll
        for (int k = 0; k < V; ++k) m[k] = floor_val;   // 0 folds a preceding ReLU into the pool
        // branch-free: a tap outside the image is replaced by the nearest tap inside, which belongs to the same window (max is
        // idempotent), so the nine loads are unconditional and all in flight; neighbouring windows share taps through L1
        float v[9][V];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int ih = min(max(2 * oh - 1 + dy, 0), H - 1);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int iw = min(max(2 * ow - 1 + dx, 0), W - 1);
                Vec<T, V>::load(img + ((long long)ih * W + iw) * C, v[dy * 3 + dx]);
            }
        }
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int k = 0; k < V; ++k) m[k] = fmaxf(m[k], v[t][k]);
        Vec<T, V>::store_stream(out + (long long)p * C, m);
    }
}

// ---------------------------------------------------------------------------------------------------
// Multi-head self-attention over the NT = 64 patch tokens of MAPTransformerEncoer (attention_pnp_net.py:126-157; timm
// 0.9.6 `Attention.forward`: softmax(q k^T * hd^-0.5) v per head).  qkv: (B, NT, 3, NH, HD) as produced by the qkv Linear,
// out: (B, NT, NH*HD).  One CTA per (head, RoI), one thread per query token: K and V of the head sit in shared memory
// (every thread walks the same row -> broadcast reads), scores / softmax / output row stay in registers, fp32 math.
// ---------------------------------------------------------------------------------------------------
template <typename T, int NT, int HD>
__global__ void __launch_bounds__(NT)
mhsa_tokens_kernel(const T *__restrict__ qkv, T *__restrict__ out, int NH, float scale) {
    __shared__ float sK[NT][HD], sV[NT][HD];
    const int h = blockIdx.x, t = threadIdx.x;
    const long long b = blockIdx.y;
    const int row = 3 * NH * HD;
    const T *base = qkv + b * NT * row + h * HD;
    float q[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) {
        q[d] = to_acc<T>(base[(long long)t * row + d]) * scale;
        sK[t][d] = to_acc<T>(base[(long long)t * row + NH * HD + d]);
        sV[t][d] = to_acc<T>(base[(long long)t * row + 2 * NH * HD + d]);
    }
    __syncthreads();
    float s[NT], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) a = fmaf(q[d], sK[j][d], a);
        s[j] = a;
        mx = fmaxf(mx, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        s[j] = __expf(s[j] - mx);
        sum += s[j];
    }
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int d = 0; d < HD; ++d) o[d] = fmaf(s[j], sV[j][d], o[d]);
    const float inv = 1.f / sum;
    T *dst = out + (b * NT + t) * (long long)(NH * HD) + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) dst[d] = from_acc<T, float>(o[d] * inv);
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
