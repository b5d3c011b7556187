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

/* releases the device scratch cached by the *_host entry points */
int gp_host_cache_release(void);

/* Tiling of the sampling kernels: output-tile height/width, groups per CTA and channels per lane for
 * 16-bit storage in the forward kernel (4 or 8); values <= 0 keep the current setting; tile sizes are rounded
 * down to powers of two; defaults 8, 8, 2, 8; also read once from
 * GP_TILE_H / GP_TILE_W / GP_GS / GP_VEC16.  Used by the tuning sweeps in tools/; results never depend on it
 * beyond floating-point summation order. */
int gp_set_tuning(int tile_h, int tile_w, int groups_per_cta, int vec16);

/* number of kernels this library has launched since load / since the last reset (bench.py's
 * `gpu_launches` is read from here, it is not an estimate) */
uint64_t gp_launch_count(void);
void gp_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* GIVEPOSE_B200_H_ */
