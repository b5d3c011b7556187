/*
 * givepose_b200 -- C ABI of the B200-native DCNv3 / PoseNet hot path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's FFI for this path
 * binds.  In the reference that FFI is the pybind11 module `DCNv3`
 *     network/ops_dcnv3/src/vision.cpp:14-17      (dcnv3_forward, dcnv3_backward)
 *     network/ops_dcnv3/src/dcnv3.h:20-59         (device dispatch, "Not implemented on the CPU")
 *     network/ops_dcnv3/src/cuda/dcnv3_cuda.cu:21-85, :87-174   (host wrappers)
 * taking at::Tensor; here the same calls take plain device pointers + a POD descriptor, so the
 * library has no torch (or Python) dependency.  `givepose_b200/dropin/DCNv3.py` is the ctypes stub
 * that re-creates the reference's Python-visible module on top of it.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t value for CUDA failures, or a
 *     negative GP_ERR_* code for argument errors; gp_error_string() explains any of them.
 *     (The reference only printf()s kernel-launch errors, dcnv3_im2col_cuda.cuh:913-916; we return them.)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Work is enqueued,
 *     never synchronised, exactly like the reference (dcnv3_cuda.cu:72,150) -- except the *_host
 *     entry points, which are synchronous by contract.
 *   - pointers are borrowed for the duration of the enqueued work; all tensors are dense,
 *     channel-last, and must be 16-byte aligned (torch allocations are 512-byte aligned).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef GIVEPOSE_B200_H_
#define GIVEPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 1

/* storage dtypes; arithmetic is fp32 for F32/BF16/F16 and fp64 for F64 (the reference's opmath_t,
 * dcnv3_im2col_cuda.cuh:30; the reference dispatches double/float/half, dcnv3_cuda.cu:69 -- BF16 is new) */
enum gp_dtype { GP_F32 = 0, GP_BF16 = 1, GP_F16 = 2, GP_F64 = 3 };

enum gp_error {
    GP_OK = 0,
    GP_ERR_NULL = -1,        /* a required pointer is NULL                                   */
    GP_ERR_SHAPE = -2,       /* C != G*gc, non-positive dims, Ho/Wo mismatch (dcnv3_cuda.cu:50-53) */
    GP_ERR_DTYPE = -3,       /* unknown gp_dtype                                              */
    GP_ERR_ALIGN = -4,       /* pointer not 16-byte aligned                                   */
    GP_ERR_WORKSPACE = -5,   /* workspace too small                                           */
    GP_ERR_UNSUPPORTED = -6  /* e.g. remove_center with an even / non-square kernel           */
};

/* Geometry of one DCNv3 core call (argument list of dcnv3_forward, src/dcnv3.h:20-26).
 * Ho/Wo must equal gp_dcnv3_out_size() of H/W (dcnv3_cuda.cu:40-45).
 *
 * offset / mask addressing is FLAT, as in the reference kernel (dcnv3_im2col_cuda.cuh:229,243-244):
 * with q = (b*Ho+oh)*Wo+ow, group g and point p = i*kh + j (kernel WIDTH index i is the slow one),
 *     offset[((q*G+g)*P + p)*2 + 0] = x (width) offset,  [... + 1] = y (height) offset
 *     mask  [ (q*G+g)*P + p ]
 * P = kh*kw - remove_center.  Buffers may be larger than that prefix (the stride-2 in-model calls pass
 * full-resolution tensors); only the prefix is read / written. */
typedef struct gp_dcnv3_desc {
    int32_t N, H, W;          /* input is (N, H, W, G*gc) channel-last                       */
    int32_t G, gc;            /* groups, channels per group                                  */
    int32_t kh, kw, sh, sw, ph, pw, dh, dw;
    int32_t remove_center;    /* 0 / 1                                                       */
    int32_t Ho, Wo;           /* output is (N, Ho, Wo, G*gc)                                 */
    float offset_scale;
} gp_dcnv3_desc;

int gp_abi_version(void);
const char *gp_error_string(int code);

/* (size + 2*pad - (dil*(k-1)+1)) / stride + 1     -- dcnv3_cuda.cu:40-45 */
int gp_dcnv3_out_size(int size, int k, int stride, int pad, int dil);

/* ---- device-pointer entry points (what a framework plugin calls) ------------------------------- */

/* replaces dcnv3_cuda_forward (dcnv3_cuda.cu:21-85) / dcnv3_im2col_cuda (cuh:890-917).
 * `mask` holds post-softmax weights.  Every element of `out` is written (no pre-zeroing needed). */
int gp_dcnv3_forward(const void *input, const void *offset, const void *mask, void *out,
                     const gp_dcnv3_desc *desc, int dtype, void *stream);

/* Same sampling, but `mask_logits` are the raw outputs of the mask Linear and the softmax over the P
 * points of each (pixel, group) is fused into the sampler (modules/dcnv3.py:332-334 + core). */
int gp_dcnv3_forward_softmax(const void *input, const void *offset, const void *mask_logits, void *out,
                             const gp_dcnv3_desc *desc, int dtype, void *stream);

/* As gp_dcnv3_forward_softmax, but offsets and mask logits come PACKED as the output rows of ONE fused Linear
 * (modules/dcnv3.py:330-334: `offset` and `mask` read the same activation): row q of `offset_mask` holds
 * [G*P*2 offsets | G*P mask logits | padding] and rows are `pitch` elements apart (pitch >= G*P*3 and even;
 * 16-byte aligned base).  Tiled kernels only: GP_ERR_UNSUPPORTED where gp_dcnv3_forward would take its
 * generic kernel (odd gc, fp64). */
int gp_dcnv3_forward_softmax_packed(const void *input, const void *offset_mask, void *out, long long pitch,
                                    const gp_dcnv3_desc *desc, int dtype, void *stream);

/* bytes of scratch gp_dcnv3_backward needs (0 for F32/F64: grads accumulate in place; for BF16/F16 an
 * fp32 image of grad_input, as the reference does for half, dcnv3_cuda.cu:126-133,168-173). */
size_t gp_dcnv3_backward_workspace(const gp_dcnv3_desc *desc, int dtype);

/* replaces dcnv3_cuda_backward (dcnv3_cuda.cu:87-174) / dcnv3_col2im_cuda (cuh:919-1094).
 * grad_input has the shape of input; grad_offset / grad_mask have grad_offset_elems / grad_mask_elems
 * elements (>= the flat prefix): the prefix is written, the tail is zero-filled by this call, and
 * grad_input is zero-filled before accumulation -- the caller passes uninitialised buffers. */
int gp_dcnv3_backward(const void *input, const void *offset, const void *mask, const void *grad_out,
                      void *grad_input, void *grad_offset, void *grad_mask, size_t grad_offset_elems,
                      size_t grad_mask_elems, void *workspace, size_t workspace_bytes,
                      const gp_dcnv3_desc *desc, int dtype, void *stream);

/* Parity hook for the integer part of the contract: for k = (q*G+g)*P + p writes
 *   hw_low[2k] = floor(loc_h), hw_low[2k+1] = floor(loc_w)   (0,0 when the sample is out of range)
 *   flags[k]   = bit0 in_range (cuh:268-269) | bit1..bit4 per-corner bounds checks (cuh:57,62,67,72)
 * computed by the SAME device function the forward/backward kernels use. */
int gp_dcnv3_sample_index(const void *offset, int32_t *hw_low, uint8_t *flags, const gp_dcnv3_desc *desc,
                          int dtype, void *stream);

/* ---- host-buffer entry points (end-to-end: H2D + kernels + D2H inside the call, synchronous) ---- */

int gp_dcnv3_forward_host(const void *h_input, const void *h_offset, const void *h_mask, void *h_out,
                          size_t offset_elems, size_t mask_elems, const gp_dcnv3_desc *desc, int dtype,
                          int device);

int gp_dcnv3_backward_host(const void *h_input, const void *h_offset, const void *h_mask,
                           const void *h_grad_out, void *h_grad_input, void *h_grad_offset,
                           void *h_grad_mask, size_t offset_elems, size_t mask_elems,
                           const gp_dcnv3_desc *desc, int dtype, int device);

/* One training-style call on host buffers: forward + backward of the same inputs, uploaded ONCE, pipelined over `chunks`
 * RoI chunks (1..64) so that H2D of chunk c+1, the kernels of chunk c and D2H of chunk c-1 overlap.  RoI chunks are exact:
 * the flat offset/mask addressing is linear in the RoI index.  Synchronous; pinned host memory is needed for overlap. */
int gp_dcnv3_forward_backward_host(const void *h_input, const void *h_offset, const void *h_mask,
                                   const void *h_grad_out, void *h_out, void *h_grad_input, void *h_grad_offset,
                                   void *h_grad_mask, size_t offset_elems, size_t mask_elems,
                                   const gp_dcnv3_desc *desc, int dtype, int device, int chunks);

/* releases the device scratch cached by the *_host entry points */
int gp_host_cache_release(void);

/* ---- PoseNet forward: fused glue kernels around the core (fp32 parameters, activations of `dtype`) -------- */

/* x1 = GELU(LayerNorm_eps(DWConv3x3_pad1(x) + bias)) of the DCNv3 module (modules/dcnv3.py:269-283,329), channel-last
 * (N,H,W,C), C in {128,256,512}.  Only the first `rows` pixels of the flat N*H*W pixel list are computed -- the core
 * reads offset/mask through their flat N*Ho*Wo-row prefix, so the stride-2 in-model calls need a quarter of x1.
 * w_t: depthwise weights transposed to [9][C] (tap-major: w_t[(ky*3+kx)*C + c] = weight[c,0,ky,kx]). */
int gp_dwconv3x3_ln_gelu(const void *x, const float *w_t, const float *bias, const float *ln_w, const float *ln_b,
                         void *out, int N, int H, int W, int C, long long rows, float eps, int dtype, void *stream);

/* First MAPEncoder layer (conv_pnp_net.py:259-272, network/dcnv3.py:32-38): DCNv3_C's 1x1 convolution lifts the K = 3
 * coordinate channels to C, and both consumers are linear in its output, so that C-channel tensor is never written:
 *   gp_small_k_linear             out(rows,C) = x(rows,K) . w_t[K][C] + bias     (w_t = (W_ip W_c)^T, bias = W_ip b_c + b_ip:
 *                                 input_proj(conv(x)), modules/dcnv3.py:325, composed by the caller in fp32)
 *   gp_smallk_dwconv3x3_ln_gelu   GELU(LN(DWConv3x3(conv(x)))) for the first `rows` pixels of x (N,H,W,K);
 *                                 w_eff[9][K+1][C]: w_eff[t][k][c] = w_dw[c,t] W_c[c,k], w_eff[t][K][c] = w_dw[c,t] b_c[c]
 *                                 (the conv bias enters only through taps inside the image: zero padding applies to conv(x)).
 * Supported: K == 3. */
int gp_small_k_linear(const void *x, const float *w_t, const float *bias, void *out, long long rows, int K, int C, int dtype,
                      void *stream);
int gp_smallk_dwconv3x3_ln_gelu(const void *x, const float *w_eff, const float *bias, const float *ln_w, const float *ln_b,
                                void *out, int N, int H, int W, int K, int C, long long rows, float eps, int dtype, void *stream);

/* The whole first-layer DCNv3 module in one kernel: out = output_proj(core(input_proj(conv1x1(x)), offset, softmax(mask)))
 * for the K = 3 channel input of MAPEncoder's first DCNv3_C (conv_pnp_net.py:259-272, network/dcnv3.py:32-38,
 * modules/dcnv3.py:318-356).  All three projections are linear and the core is linear in its input, so the kernel samples the
 * K-channel map itself (K values + the sum of the in-image weights per (pixel, group)) and applies ONE composed
 * [G*(K+1)] -> C_out map: w2[(g*(K+1)+j)*C_out + o] = sum_c Wo[o, g*gc+c] * Wp[g*gc+c, j] (j < K), j == K: the same with bp
 * (Wp, bp = input_proj o conv1x1), bias = output_proj bias; composed by the caller in fp64, stored fp32.
 * x (N,H,W,K), offset / mask_logits: the flat [N*Ho*Wo*G*P] prefix is read (cuh:243-244), softmax over P fused;
 * out (N,Ho,Wo,C_out).  Indices / bounds are those of gp_dcnv3_forward (same locate()).  Supported: K 3, G 4, 3x3, C_out 256. */
int gp_dcnv3_smallk_fused(const void *x, const void *offset, const void *mask_logits, const float *w2, const float *bias, void *out,
                          size_t offset_elems, size_t mask_elems, const gp_dcnv3_desc *d, int K, int C_out, int dtype, void *stream);

/* y = act(GroupNorm_G(x)) on channel-last (N,H,W,C) activations (layer_utils.py:32-60 "GN", conv_module.py order
 * conv -> norm -> act); act: 0 none, 1 ReLU, 2 GELU.  stats: scratch of at least gp_groupnorm_workspace_floats(N,H,W,G)
 * floats (no need to clear it).  Sums are taken in a fixed order (no atomics): results are bit-reproducible. */
size_t gp_groupnorm_workspace_floats(int N, int H, int W, int G);
int gp_groupnorm_act(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta, int N,
                     int H, int W, int C, int G, float eps, int act, int dtype, void *stream);

/* Backward of gp_groupnorm_act for the training step (replaces torch's native_group_norm_backward + gelu/threshold
 * backward that follow `get_norm("GN")` / `get_nn_act_func` under autograd, layer_utils.py:32-94).  stats = the first N*G*2
 * floats gp_groupnorm_act left in its scratch ((mean, rstd) per (n, group)); x is the forward INPUT, dy the gradient of the
 * activation output (same dtype / layout as x); dx gets the input gradient, dgamma / dbeta [C] fp32 are OVERWRITTEN.
 * ws: scratch of gp_groupnorm_backward_workspace_floats floats.  No atomics: bit-reproducible. */
size_t gp_groupnorm_backward_workspace_floats(int N, int H, int W, int C, int G);
int gp_groupnorm_act_backward(const void *x, const void *dy, const float *stats, const float *gamma, const float *beta, void *dx,
                              float *dgamma, float *dbeta, float *ws, size_t ws_floats, int N, int H, int W, int C, int G, int act,
                              int dtype, void *stream);

/* y = Conv1x1_{C->OC}(act(GroupNorm_G(x))) + bias: the decoder's last ConvModule norm/activation fused with its
 * out_layer (xyz_head.py:349-366, Conv1x1 256 -> 3); the normalised C-channel activation is never written.
 * x (N,H,W,C) channel-last, y (N,H,W,OC) of `dtype`; w [OC][C] and bias [OC] fp32.  Supported: C == 256, OC == 3.
 * For 16-bit storage GELU is evaluated with a tanh-form fit of the erf GELU (error < 2.5e-4 |x|, below the bf16
 * rounding of the result); fp32 keeps the erf form (abs error 1.5e-7).  The same holds for gp_groupnorm_act. */
int gp_groupnorm_act_conv1x1(const void *x, void *y, float *stats, size_t stats_floats, const float *gamma, const float *beta,
                             const float *w, const float *bias, int N, int H, int W, int C, int G, float eps, int act, int OC,
                             int dtype, void *stream);

/* Apply pass of gp_groupnorm_act / gp_groupnorm_act_conv1x1 with (mean, rstd) per (n, group) already in stats[0 .. N*G*2)
 * (stats_floats >= N*G*2), e.g. from gp_conv3x3_gn_bf16 + gp_groupnorm_finalize: the statistics pass over x is skipped. */
int gp_groupnorm_apply(const void *x, void *y, const float *stats, size_t stats_floats, const float *gamma, const float *beta, int N,
                       int H, int W, int C, int G, float eps, int act, int dtype, void *stream);
int gp_groupnorm_apply_conv1x1(const void *x, void *y, const float *stats, size_t stats_floats, const float *gamma, const float *beta,
                               const float *w, const float *bias, int N, int H, int W, int C, int G, float eps, int act, int OC,
                               int dtype, void *stream);
/* stats[n][g] = (mean, rstd) from per-slab partial (sum, sum of squares) pairs laid out [N][slabs][G][2] (fp32), summed in slab
 * order (bit-reproducible); count = elements per (n, group) = H*W*C/G. */
int gp_groupnorm_finalize(const float *partial, float *stats, int N, int G, int slabs, long long count, float eps, void *stream);

/* The coordinate-map decoder's 3x3 convolutions (network/xyz_head.py:195-366 -> ConvModule, network/torch_utils/layers/
 * conv_module.py:57-234: Conv2d(Cin, 256, 3, stride 1, padding 1, bias=False)) as a hand-written tcgen05 implicit GEMM with the
 * GroupNorm(32) statistics of the result produced in the epilogue (conv3x3_tc.cu).
 *   x (N,H,W,Cin) bf16 channel-last; w_packed [256][3][3][Cin] bf16 (= conv.weight.permute(0,2,3,1)); y (N,H,W,256) bf16.
 *   partial: NULL, or N * gp_conv3x3_gn_slabs(H,W) * 32 * 2 floats that receive the (sum, sum of squares) of the fp32
 *   accumulators per (image, 64-pixel slab, group of 8 channels) in the layout gp_groupnorm_finalize reads.
 * Supported: Cout == 256, Cin % 64 == 0, W in {8, 16, 32, 64} with H a multiple of 256/W (whole 256-pixel row blocks).
 * fp32 accumulation over K = 9*Cin in tensor memory; zero padding comes from TMA's out-of-bounds fill. */
size_t gp_conv3x3_gn_slabs(int H, int W);
/* The same convolution with the PRODUCER layer's GroupNorm(32) + GELU applied to its operand on the way to the tensor core
 * (ConvModule -> ConvModule at one resolution, conv_module.py order conv -> norm -> act): x_raw is the raw (un-normalised) output
 * of the previous convolution, in_stats its (mean, rstd) pairs [N][32][2] (gp_groupnorm_finalize), in_gamma / in_beta [Cin] fp32.
 * Bit-identical to gp_groupnorm_apply(GELU) followed by gp_conv3x3_gn_bf16, without the apply pass over the activation.
 * CTA-pair kernel only (gp_conv3x3_set_pair(1), the default); Cin a multiple of 256. */
int gp_conv3x3_gn_bf16_fused_in(const void *x_raw, const float *in_stats, const float *in_gamma, const float *in_beta,
                                const void *w_packed, void *y, float *partial, int N, int H, int W, int Cin, int Cout, void *stream);
/* Kernel variant: 0 = one CTA per 256 x 256 tile (two accumulators = all of TMEM), 1 = CTA pair on one TPC
 * (tcgen05.mma.cta_group::2 M256: half of every weight block per SM, two accumulator sets so the epilogue overlaps the next tile).
 * Returns the previous setting; the initial value comes from GP_CONV_PAIR in the environment.  gp_conv3x3_gn_slabs follows it. */
int gp_conv3x3_set_pair(int on);
int gp_conv3x3_gn_bf16(const void *x, const void *w_packed, void *y, float *partial, int N, int H, int W, int Cin, int Cout,
                       void *stream);

/* ---- RoI input pipeline in front of PoseNet.forward -------------------------------------------
 * Replaces the per-RoI host work of the reference's loaders (evaluation/load_data_eval.py:256-289,
 * datasets/load_data_nocs.py:277-305): crop_resize_by_warp_affine (tools/dataset_utils.py:101-114 = get_affine_transform
 * :116-157 + cv2.warpAffine INTER_NEAREST), the (x/255 - mean)/std normalisation + HWC->CHW, and the crop of
 * get_2d_coord_np (:8-30).  Index work: bit-exact against OpenCV 4.8's warpAffine (the reference's pin).
 *
 * gp_roi_affine_inverse (HOST, synchronous): for each RoI the 2x3 matrix cv2.getAffineTransform returns for the reference's
 *   three float32 point pairs, inverted the way cv::warpAffine inverts it; center (B,2), scale (B,) double (bbox_center,
 *   img_scale of the loaders); minv (B,6) double, dst -> src.
 * gp_roi_crop: images (n_images,H,W,3) uint8 RGB, masks (n_masks,H,W) uint8, image_index / mask_index / inst_id (B,) int32
 *   (inst_id < 0: roi_mask = float(mask); else roi_mask = (mask == inst_id)), minv_img / minv_out (B,6) double DEVICE pointers for
 *   the img_size and out_res outputs, lut [3][256] fp32 = float((v/255.0 - mean[c])/std[c]) computed in double;
 *   roi_img (B,3,S,S), roi_mask (B,1,S,S), roi_coord_2d (B,2,R,R) fp32 -- any of the three may be NULL.  B <= 65535. */
int gp_roi_affine_inverse(const double *center, const double *scale, int B, int out_size, double *minv);
int gp_roi_crop(const uint8_t *images, int n_images, int H, int W, const int *image_index, const uint8_t *masks, int n_masks,
                const int *mask_index, const int *inst_id, const double *minv_img, const double *minv_out, const float *lut,
                float *roi_img, float *roi_mask, float *roi_coord_2d, int B, int img_size, int out_res, void *stream);

/* full_img of the loaders with FLAGS.resize_full (the default): cv2.resize(frame, (dw, dh)) -- INTER_LINEAR, 8-bit fixed point --
 * then (v/255 - mean)/std through the same table, HWC -> CHW (evaluation/load_data_eval.py:336-338).  images (n_images,H,W,3)
 * uint8, image_index (B,) int32 or NULL (RoI b reads frame b), out (B,3,dh,dw) fp32.  Bit-exact against OpenCV 4.8 for
 * H >= dh and W >= dw; upscaling is refused (GP_ERR_UNSUPPORTED). */
int gp_resize_linear_u8_normalize(const uint8_t *images, int n_images, int H, int W, const int *image_index, const float *lut, float *out,
                                  int B, int dh, int dw, void *stream);

/* Dense layer on the tcgen05 tensor cores: y[M,N] = act(x[M,K] . w[N,K]^T + bias[N]), x / w / y bf16 row-major, bias fp32,
 * fp32 accumulation in TMEM, bias + activation in the epilogue (act: 0 none, 1 LeakyReLU(slope), 2 ReLU).  The shape of
 * every Linear / 1x1 convolution of the heads: DCNv3 input_proj / output_proj / offset / mask (modules/dcnv3.py:325-354),
 * feat_reducer (PoseNet.py:158), fc1||fc1_z, fc2, fc2_z with their LeakyReLU(0.1) (conv_pnp_net.py:172-199).
 * K % 8 == 0 (16-byte row pitch); any M, N (tails are masked).  TMA operand loads, 3-stage mbarrier ring, two CTAs per SM. */
int gp_linear_bf16(const void *x, const void *w, const float *bias, void *y, int M, int N, int K, int act, float slope,
                   void *stream);
/* Kernel choice of gp_linear_bf16 for large shapes: 2 (default) = CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles, two accumulator
 * sets) when at least 64 such tiles exist and K >= 512, 0 = always one CTA per tile, 1 = CTA pairs whenever legal (N % 8 == 0,
 * M, N >= 128).  Returns the previous mode; GP_LINEAR_PAIR in the environment sets the initial one. */
int gp_linear_set_pair(int mode);

/* Multi-head self-attention over the 64 patch tokens of MAPTransformerEncoer (attention_pnp_net.py:126-157, the
 * `--nocsmap_encoder=att` alternative to MAPEncoder; timm 0.9.6 Attention.forward): out = softmax(q k^T * scale) v per head.
 * qkv (B, NT, 3, NH, HD) as the qkv Linear emits it, out (B, NT, NH*HD), both of `dtype`.  Supported: NT == 64, HD == 32. */
int gp_mhsa_tokens(const void *qkv, void *out, int B, int NT, int NH, int HD, float scale, int dtype, void *stream);

/* Operand packing for a 7x7 / stride-2 / pad-3 stem convolution over 3 input channels (network/resnet.py:104, the
 * stand-in backbone of the synthetic runs) evaluated as a 4x4 / stride-1 convolution over the 2x2 space-to-depth image:
 * img fp32 (N,3,H,W) NCHW  ->  out (N, H/2+3, W/2+3, 16) channel-last of `dtype`,
 * out[n, 2+y2, 2+x2, c*4 + ry*2 + rx] = img[n, c, 2*y2+ry, 2*x2+rx], zero elsewhere (spatial padding 2 before / 1 after,
 * channels 12..15). */
int gp_stem_s2d_pack(const float *img, void *out, int N, int H, int W, int dtype, void *stream);

/* nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True) of TopDownXyzHead (xyz_head.py:262-265) on channel-last
 * activations: (N,H,W,C) -> (N,2H,2W,C). */
int gp_upsample_bilinear2x(const void *x, void *y, int N, int H, int W, int C, int dtype, void *stream);
/* Its backward as a deterministic gather: dy (N,2H,2W,C) -> dx (N,H,W,C), the exact transpose of the forward weights. */
int gp_upsample_bilinear2x_backward(const void *dy, void *dx, int N, int H, int W, int C, int dtype, void *stream);

/* MaxPool2d(3, stride 2, pad 1) on channel-last activations (stem of the stand-in ResNet backbone, resnet.py:106):
 * (N,H,W,C) -> (N,(H-1)/2+1,(W-1)/2+1,C).  relu != 0 computes maxpool(relu(x)) (= relu(maxpool(x))) in the same pass. */
/* y = relu(y + bias[c] + residual) in place on channel-last (rows, C) activations: tail of a residual block of the stand-in
 * backbone when the convolution runs without cuDNN's fused add+ReLU epilogue.  bias fp32 [C]; C % 8 == 0. */
int gp_bias_add_relu(void *y, const void *residual, const float *bias, long long rows, int C, int dtype, void *stream);
/* Stem of the stand-in backbone on the tcgen05 tensor cores: y (N, Hp-3, Wp-3, 64) = relu(conv4x4/1(packed) + bias), packed =
 * gp_stem_s2d_pack's (N, Hp, Wp, 16) bf16 image, w [64][4*4*16] bf16 (tap-major: (dy, dx, c)), bias fp32 [64].  TMA does the
 * im2col through a tensor map with overlapping rows; one new 16 KB block per output row, weights resident in shared memory.
 * pool != 0: the 3x3/2 max-pool (pad 1) that follows in the backbone is applied in the epilogue (a ring of the last three
 * activated rows in shared memory) and y is the pooled (N, 64, 64, 64) tensor: the pre-pool activation never leaves the SM.
 * Supported: Wp - 3 == 128 (256 x 256 crops), bf16; anything else returns GP_ERR_UNSUPPORTED (callers fall back to cuDNN). */
int gp_stem_s2d_gemm(const void *packed, const void *w, const float *bias, void *y, int N, int Hp, int Wp, int pool, void *stream);
int gp_maxpool3x3s2(const void *x, void *y, int N, int H, int W, int C, int relu, int dtype, void *stream);

/* rot6 (B,6) + t (B,3: centroid dx, dy, relative z) -> ego rotation (B,3,3) and translation (B,3):
 * rot6d_to_mat_batch (rot_reps.py:34-55), back-projection (pose_from_pred_centroid_z.py:78-119, z_type REL,
 * z_calib = fx/590 for wild6d else 1) and the allocentric->egocentric rotation (pose_utils/utils.py:29-60) that the
 * reference runs in a per-RoI host loop (:139-157).  cam: (B,3,3) if cam_batched else (3,3). */
int gp_pose_decode(const float *rot6, const float *t, const float *cam, int cam_batched, const float *centers,
                   const float *whs, const float *ratios, float *rot_out, float *trans_out, int B, int is_allo,
                   float z_calib, void *stream);

/* Tiling of the sampling kernels: output-tile height/width, groups per CTA and channels per lane for
 * 16-bit storage in the forward kernel (4 or 8); values <= 0 keep the current setting; tile sizes are rounded
 * down to powers of two; defaults 8, 8, 2, 8; also read once from
 * GP_TILE_H / GP_TILE_W / GP_GS / GP_VEC16.  Used by the tuning sweeps in tools/; results never depend on it
 * beyond floating-point summation order. */
int gp_set_tuning(int tile_h, int tile_w, int groups_per_cta, int vec16);

/* Further kernel-selection knobs (tuning sweeps and A/B measurements only; results never depend on them beyond
 * floating-point summation order).  Returns GP_ERR_SHAPE for an unknown key / value.
 *   GP_OPT_BWD_MODE      0 = one-pass scatter backward (default): one 16-byte reduction per lane per corner
 *                        1 = split backward: grad_offset/grad_mask kernel + grad_input with in-SM pre-aggregation
 *                            (dcnv3_gin_binned: counting sort by footprint cell, register accumulation, one reduction per
 *                            destination line of the tile's window: 7.6x fewer reduction sectors, but the sort + gather
 *                            costs more SM issue slots than the reductions cost L2 bandwidth -- measured 1.86 ms against
 *                            1.38 ms at BASELINE config 2, profiles/r02_it1_sweep_bwd_split_vs_scatter.json)
 *                        2 = ONE kernel (dcnv3_bwd_fused): the same gathers, and grad_input aggregated inside the SM by a
 *                            counting sort of the CTA's samples by footprint cell + a register row walk (each sample
 *                            visited once, ~3x fewer reductions), so the two phases overlap across the CTAs of an SM
 *   GP_OPT_GIN_TILE_H/W  output tile of the binned grad_input kernel (powers of two, default 8 x 8)
 *   GP_OPT_GIN_THREADS   its CTA size: 128, 192 (default) or 256
 *   GP_OPT_FWD_MODE      0 = dcnv3_fwd_tile (per-thread offset / mask row reads, 24-byte records), 1 = dcnv3_fwd_rows for 3x3
 *                        kernels (default): rows staged by TMA, 16-byte records; other shapes always take mode 0 */
enum gp_option { GP_OPT_BWD_MODE = 0, GP_OPT_GIN_TILE_H = 1, GP_OPT_GIN_TILE_W = 2, GP_OPT_GIN_THREADS = 3,
                 GP_OPT_FWD_MODE = 4, GP_OPT_LAST_FWD_KERNEL = 5 /* read-only: which kernel the last forward call launched:
                 0 dcnv3_fwd_tile, 1 dcnv3_fwd_rows, 2 dcnv3_fwd_generic, -1 none yet */ };
int gp_set_option(int key, int value);
int gp_get_option(int key);

/* number of kernels this library has launched since load / since the last reset (bench.py's
 * `gpu_launches` is read from here, it is not an estimate) */
uint64_t gp_launch_count(void);
void gp_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* GIVEPOSE_B200_H_ */
