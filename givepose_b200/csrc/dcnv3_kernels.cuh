// givepose_b200 -- DCNv3 deformable-sampling kernels for sm_100a (forward, backward, index hook).
//
// Replaces the reference device code network/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh
//   forward  :216-282 (one thread per output SCALAR, offsets/mask re-read per channel, scalar gathers)
//   backward :285-888 (block = group_channels threads, 9x block-wide smem reductions, scalar atomics)
// with a different decomposition (nothing is translated):
//
//   unit      = one (output pixel q, group g): P sampling points, gc channels.
//   CTA       = a tile_h x tile_w patch of output pixels x `gs` groups of ONE image, so the input rows it
//               gathers form a compact window that stays in L1 (channel-last rows of one group are
//               contiguous: gc*sizeof(T) bytes = one 128-byte line for gc=32 fp32).
//   staging   = the CTA's offset / mask rows are read once, coalesced, widened to fp32 (and, for the
//               fused variant, soft-maxed over the P points) into shared memory; the sampling loop then
//               reads them with conflict-free broadcast LDS.
//   thread    = VEC channels (16 bytes) of one unit; L = gc/VEC consecutive lanes form a unit, so every
//               corner gather of a unit is one fully-used 128-bit-per-lane coalesced request.
//   backward  = same tiling; grad_input goes out as 16-byte vector reductions (REDG.ADD.F32x4) that
//               resolve in L2 on the tile's window; the three per-point sums over the gc channels are
//               butterfly-reduced with warp shuffles over the L lanes (no block barriers), parked in the
//               staging buffer and written back coalesced.
#pragma once

#include "gp_common.cuh"

namespace gp {

constexpr int kTileThreads = 256;

// ---------------------------------------------------------------------------------------------------
// CTA decode + staging shared by forward and backward
// ---------------------------------------------------------------------------------------------------
struct TileCtx {
    int b, oh0, ow0, g0, TP, n_ul;
};

__device__ __forceinline__ TileCtx decode_tile(const KParams &p) {
    TileCtx t;
    int bid = blockIdx.x;
    const int gch = bid % p.gchunks;
    bid /= p.gchunks;
    const int tx = bid % p.tiles_x;
    bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    t.b = bid / p.tiles_y;
    t.oh0 = ty * p.tile_h;
    t.ow0 = tx * p.tile_w;
    t.g0 = gch * p.gs;
    t.TP = p.tile_h * p.tile_w;
    t.n_ul = t.TP * p.gs;
    return t;
}

// smem layout: s_off[(ul*P + pt)*2 + {0,1}], s_msk[ul*P + pt], ul = g_local*TP + pix  (group-major so that
// consecutive passes of the sampling loop work on one group => one L1-resident window at a time)
template <typename T, bool SOFTMAX>
__device__ __forceinline__ void stage_offsets_mask(const T *__restrict__ off, const T *__restrict__ msk, float *s_off,
                                                   float *s_msk, const KParams &p, const TileCtx &t) {
    const int P = p.P;
    const int row2 = p.gs * P * 2, row1 = p.gs * P;
    for (int e = threadIdx.x; e < t.TP * row2; e += blockDim.x) {
        const int pix = e / row2, r = e - pix * row2;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        float v = 0.f;
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            v = to_acc<T>(__ldg(off + (q * p.G + t.g0) * (long long)(P * 2) + r));
        }
        const int gl = r / (P * 2);
        s_off[(gl * t.TP + pix) * (P * 2) + (r - gl * P * 2)] = v;
    }
    for (int e = threadIdx.x; e < t.TP * row1; e += blockDim.x) {
        const int pix = e / row1, r = e - pix * row1;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        float v = 0.f;
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            v = to_acc<T>(__ldg(msk + (q * p.G + t.g0) * (long long)P + r));
        }
        const int gl = r / P;
        s_msk[(gl * t.TP + pix) * P + (r - gl * P)] = v;
    }
    __syncthreads();
    if (SOFTMAX) {   // softmax over the P logits of each (pixel, group) row: modules/dcnv3.py:332-333
        for (int ul = threadIdx.x; ul < t.n_ul; ul += blockDim.x) {
            float *row = s_msk + ul * P;
            float mx = row[0];
            for (int i = 1; i < P; ++i) mx = fmaxf(mx, row[i]);
            float sum = 0.f;
            for (int i = 0; i < P; ++i) {
                const float e = expf(row[i] - mx);
                row[i] = e;
                sum += e;
            }
            const float inv = 1.f / sum;
            for (int i = 0; i < P; ++i) row[i] *= inv;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward, tiled + vectorised.  K3 = 3x3 kernel without remove_center (fully unrolled).
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC, int L, bool K3, bool SOFTMAX>
__global__ void __launch_bounds__(kTileThreads)
dcnv3_fwd_tile(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk, T *__restrict__ out,
               const __grid_constant__ KParams p) {
    extern __shared__ float smem[];
    const TileCtx t = decode_tile(p);
    const int P = p.P;
    float *s_off = smem, *s_msk = smem + t.n_ul * P * 2;
    stage_offsets_mask<T, SOFTMAX>(off, msk, s_off, s_msk, p, t);

    const int cl = threadIdx.x % L;
    const int C = p.C, WC = p.W * C;
    const T *in_b = in + (long long)t.b * p.H * WC + cl * VEC;

    for (int ul = threadIdx.x / L; ul < t.n_ul; ul += kTileThreads / L) {
        const int gl = ul / t.TP, pix = ul - gl * t.TP;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        if (oh >= p.Ho || ow >= p.Wo) continue;
        const int g = t.g0 + gl;
        const float p0_h_ = origin<float>(p.base_h + oh * p.sh, p.half_h, p.scale);
        const float p0_w_ = origin<float>(p.base_w + ow * p.sw, p.half_w, p.scale);
        const T *in_g = in_b + g * p.gc;
        const float2 *so = reinterpret_cast<const float2 *>(s_off) + ul * P;
        const float *sm = s_msk + ul * P;

        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;

        auto sample = [&](int i, int j, int pt_idx) {
            const float2 o = so[pt_idx];
            const float m = sm[pt_idx];
            Point<float> pt;
            locate<float>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, o.x, o.y, p.scale, p.H, p.W);
            if (pt.flags & F_IN) {
                const int base = pt.h_low * WC + pt.w_low * C;
                float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
                for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
                if (pt.flags & F_C1) Vec<T, VEC>::load(in_g + base, v1);
                if (pt.flags & F_C2) Vec<T, VEC>::load(in_g + base + C, v2);
                if (pt.flags & F_C3) Vec<T, VEC>::load(in_g + base + WC, v3);
                if (pt.flags & F_C4) Vec<T, VEC>::load(in_g + base + WC + C, v4);
                const float w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
#pragma unroll
                for (int c = 0; c < VEC; ++c) {
                    const float val = (w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c]);   // cuh:78
                    acc[c] += val * m;                                                        // cuh:270-273
                }
            }
        };

        if (K3) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) sample(i, j, i * 3 + j);
        } else {
            const int ch = p.kh / 2, cw = p.kw / 2;
            int pt_idx = 0;
            for (int i = 0; i < p.kw; ++i)
                for (int j = 0; j < p.kh; ++j)
                    if (i != cw || j != ch || !p.remove_center) sample(i, j, pt_idx++);
        }
        const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
        Vec<T, VEC>::store_stream(out + q * C + g * p.gc + cl * VEC, acc);
    }
}

// ---------------------------------------------------------------------------------------------------
// Forward, generic: any gc, any dtype (incl. f64).  One thread per output scalar.
// ---------------------------------------------------------------------------------------------------
template <typename T, bool SOFTMAX>
__global__ void __launch_bounds__(256)
dcnv3_fwd_generic(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk, T *__restrict__ out,
                  const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    const long long total = p.n_units * p.gc;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % p.gc);
        const long long unit = idx / p.gc;
        const int g = (int)(unit % p.G);
        const long long q = unit / p.G;
        const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
        const int b = (int)(q / ((long long)p.Wo * p.Ho));
        const A scale = (A)p.scale;
        const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
        const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
        const long long WC = (long long)p.W * p.C;
        const T *im = in + (long long)b * p.H * WC + g * p.gc + c;
        const T *o = off + unit * (p.P * 2);
        const T *m = msk + unit * p.P;
        A mx = (A)0, inv = (A)1;
        if (SOFTMAX) {
            mx = to_acc<T>(m[0]);
            for (int i = 1; i < p.P; ++i) mx = max(mx, to_acc<T>(m[i]));
            A s = (A)0;
            for (int i = 0; i < p.P; ++i) s += exp(to_acc<T>(m[i]) - mx);
            inv = (A)1 / s;
        }
        const int ch = p.kh / 2, cw = p.kw / 2;
        A col = (A)0;
        int k = 0;
        for (int i = 0; i < p.kw; ++i)
            for (int j = 0; j < p.kh; ++j)
                if (i != cw || j != ch || !p.remove_center) {
                    Point<A> pt;
                    locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(o[2 * k]), to_acc<T>(o[2 * k + 1]), scale,
                              p.H, p.W);
                    A w = to_acc<T>(m[k]);
                    if (SOFTMAX) w = exp(w - mx) * inv;
                    if (pt.flags & F_IN) {
                        const long long base = (long long)pt.h_low * WC + (long long)pt.w_low * p.C;
                        const A v1 = (pt.flags & F_C1) ? to_acc<T>(im[base]) : (A)0;
                        const A v2 = (pt.flags & F_C2) ? to_acc<T>(im[base + p.C]) : (A)0;
                        const A v3 = (pt.flags & F_C3) ? to_acc<T>(im[base + WC]) : (A)0;
                        const A v4 = (pt.flags & F_C4) ? to_acc<T>(im[base + WC + p.C]) : (A)0;
                        const A w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                        col += (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4) * w;
                    }
                    ++k;
                }
        out[idx] = from_acc<T, A>(col);
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward, tiled + vectorised.  gin accumulates in fp32 (grad_input itself for T=float, the workspace
// for 16-bit storage -- the reference does the same for half, dcnv3_cuda.cu:126-133).
// ---------------------------------------------------------------------------------------------------
template <typename T, int VEC, int L, bool K3>
__global__ void __launch_bounds__(kTileThreads)
dcnv3_bwd_tile(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk,
               const T *__restrict__ gout, float *__restrict__ gin, T *__restrict__ goff, T *__restrict__ gmsk,
               const __grid_constant__ KParams p) {
    extern __shared__ float smem[];
    const TileCtx t = decode_tile(p);
    const int P = p.P;
    float *s_off = smem, *s_msk = smem + t.n_ul * P * 2;
    stage_offsets_mask<T, false>(off, msk, s_off, s_msk, p, t);

    const int cl = threadIdx.x % L;
    const int C = p.C, WC = p.W * C;
    const long long img = (long long)t.b * p.H * WC;
    const T *in_b = in + img + cl * VEC;
    float *gin_b = gin + img + cl * VEC;
    const int n_pass = (t.n_ul + kTileThreads / L - 1) / (kTileThreads / L);

    for (int pass = 0; pass < n_pass; ++pass) {
        const int ul = pass * (kTileThreads / L) + threadIdx.x / L;
        const int gl = ul / t.TP, pix = ul - gl * t.TP;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        const bool valid = ul < t.n_ul && oh < p.Ho && ow < p.Wo;
        if (!__any_sync(0xffffffffu, valid)) continue;   // warp-uniform
        const int g = t.g0 + (valid ? gl : 0);
        const float p0_h_ = origin<float>(p.base_h + oh * p.sh, p.half_h, p.scale);
        const float p0_w_ = origin<float>(p.base_w + ow * p.sw, p.half_w, p.scale);
        const T *in_g = in_b + g * p.gc;
        float *gin_g = gin_b + g * p.gc;
        const int ulc = valid ? ul : 0;
        float2 *so = reinterpret_cast<float2 *>(s_off) + ulc * P;
        float *sm = s_msk + ulc * P;
        const int Hv = valid ? p.H : 0;   // invalid lanes: every sample out of range => no memory traffic

        float go[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) go[c] = 0.f;
        if (valid) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            Vec<T, VEC>::load_stream(gout + q * C + g * p.gc + cl * VEC, go);
        }

        auto sample = [&](int i, int j, int pt_idx) {
            const float2 o = so[pt_idx];
            const float m = sm[pt_idx];
            Point<float> pt;
            locate<float>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, o.x, o.y, p.scale, Hv, p.W);
            float s_m = 0.f, s_w = 0.f, s_h = 0.f;
            if (pt.flags & F_IN) {
                const int base = pt.h_low * WC + pt.w_low * C;
                float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
                for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
                if (pt.flags & F_C1) Vec<T, VEC>::load(in_g + base, v1);
                if (pt.flags & F_C2) Vec<T, VEC>::load(in_g + base + C, v2);
                if (pt.flags & F_C3) Vec<T, VEC>::load(in_g + base + WC, v3);
                if (pt.flags & F_C4) Vec<T, VEC>::load(in_g + base + WC + C, v4);
                const float w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                float tg[VEC];   // top_grad * mask, cuh:107
#pragma unroll
                for (int c = 0; c < VEC; ++c) {
                    tg[c] = go[c] * m;
                    const float val = (w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c]);
                    // grad_w_weight / grad_h_weight, cuh:114-139
                    const float gw = pt.hh * (v2[c] - v1[c]) + pt.lh * (v4[c] - v3[c]);
                    const float gh = pt.hw * (v3[c] - v1[c]) + pt.lw * (v4[c] - v2[c]);
                    s_m += go[c] * val;       // cuh:144
                    s_w += gw * tg[c];        // cuh:145 (offset_scale applied after the channel sum)
                    s_h += gh * tg[c];        // cuh:146
                }
                if (!(p.debug & 1))
#pragma unroll
                for (int c4 = 0; c4 < VEC; c4 += 4) {   // cuh:116-140: one 16-byte reduction per corner
                    if (pt.flags & F_C1) red_add_v4(gin_g + base + c4, w1 * tg[c4], w1 * tg[c4 + 1], w1 * tg[c4 + 2], w1 * tg[c4 + 3]);
                    if (pt.flags & F_C2) red_add_v4(gin_g + base + C + c4, w2 * tg[c4], w2 * tg[c4 + 1], w2 * tg[c4 + 2], w2 * tg[c4 + 3]);
                    if (pt.flags & F_C3) red_add_v4(gin_g + base + WC + c4, w3 * tg[c4], w3 * tg[c4 + 1], w3 * tg[c4 + 2], w3 * tg[c4 + 3]);
                    if (pt.flags & F_C4) red_add_v4(gin_g + base + WC + C + c4, w4 * tg[c4], w4 * tg[c4 + 1], w4 * tg[c4 + 2], w4 * tg[c4 + 3]);
                }
            }
            // sum over the gc channels of the group = butterfly over the L lanes of this unit
            if (!(p.debug & 2))
#pragma unroll
            for (int o2 = L / 2; o2 > 0; o2 >>= 1) {
                s_m += __shfl_xor_sync(0xffffffffu, s_m, o2);
                s_w += __shfl_xor_sync(0xffffffffu, s_w, o2);
                s_h += __shfl_xor_sync(0xffffffffu, s_h, o2);
            }
            if (cl == 0 && valid) {   // park the results in the (now consumed) staging slots
                so[pt_idx] = make_float2(p.scale * s_w, p.scale * s_h);
                sm[pt_idx] = s_m;
            }
        };

        if (K3) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) sample(i, j, i * 3 + j);
        } else {
            const int ch = p.kh / 2, cw = p.kw / 2;
            int pt_idx = 0;
            for (int i = 0; i < p.kw; ++i)
                for (int j = 0; j < p.kh; ++j)
                    if (i != cw || j != ch || !p.remove_center) sample(i, j, pt_idx++);
        }
    }
    __syncthreads();

    // coalesced write-back of grad_offset / grad_mask (inverse of the staging permutation)
    const int row2 = p.gs * P * 2, row1 = p.gs * P;
    for (int e = threadIdx.x; e < t.TP * row2; e += blockDim.x) {
        const int pix = e / row2, r = e - pix * row2;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            const int gl = r / (P * 2);
            goff[(q * p.G + t.g0) * (long long)(P * 2) + r] =
                from_acc<T, float>(s_off[(gl * t.TP + pix) * (P * 2) + (r - gl * P * 2)]);
        }
    }
    for (int e = threadIdx.x; e < t.TP * row1; e += blockDim.x) {
        const int pix = e / row1, r = e - pix * row1;
        const int oh = t.oh0 + pix / p.tile_w, ow = t.ow0 + pix % p.tile_w;
        if (oh < p.Ho && ow < p.Wo) {
            const long long q = ((long long)t.b * p.Ho + oh) * p.Wo + ow;
            const int gl = r / P;
            gmsk[(q * p.G + t.g0) * (long long)P + r] = from_acc<T, float>(s_msk[(gl * t.TP + pix) * P + (r - gl * P)]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Backward, generic: any gc / dtype.  One CTA (64 threads) walks units; threads stride over channels.
// GA = accumulate type of grad_input (float workspace for 16-bit storage).
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64)
dcnv3_bwd_generic(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ msk,
                  const T *__restrict__ gout, typename AccOf<T>::type *__restrict__ gin, T *__restrict__ goff,
                  T *__restrict__ gmsk, const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    __shared__ A red[2][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long WC = (long long)p.W * p.C;
    for (long long unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const int g = (int)(unit % p.G);
        const long long q = unit / p.G;
        const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
        const int b = (int)(q / ((long long)p.Wo * p.Ho));
        const A scale = (A)p.scale;
        const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
        const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
        const long long img = (long long)b * p.H * WC + g * p.gc;
        const T *o = off + unit * (p.P * 2);
        const T *m = msk + unit * p.P;
        const T *go = gout + q * p.C + g * p.gc;
        const int ch = p.kh / 2, cw = p.kw / 2;
        int k = 0;
        for (int i = 0; i < p.kw; ++i)
            for (int j = 0; j < p.kh; ++j)
                if (i != cw || j != ch || !p.remove_center) {
                    Point<A> pt;
                    locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(o[2 * k]), to_acc<T>(o[2 * k + 1]), scale,
                              p.H, p.W);
                    const A w = to_acc<T>(m[k]);
                    A s_m = (A)0, s_w = (A)0, s_h = (A)0;
                    if (pt.flags & F_IN) {
                        const long long base = img + (long long)pt.h_low * WC + (long long)pt.w_low * p.C;
                        const A w1 = pt.hh * pt.hw, w2 = pt.hh * pt.lw, w3 = pt.lh * pt.hw, w4 = pt.lh * pt.lw;
                        for (int c = threadIdx.x; c < p.gc; c += blockDim.x) {
                            const A tgrad = to_acc<T>(go[c]);
                            const A tg = tgrad * w;
                            A v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                            if (pt.flags & F_C1) { v1 = to_acc<T>(in[base + c]); atomicAdd(gin + base + c, w1 * tg); }
                            if (pt.flags & F_C2) { v2 = to_acc<T>(in[base + p.C + c]); atomicAdd(gin + base + p.C + c, w2 * tg); }
                            if (pt.flags & F_C3) { v3 = to_acc<T>(in[base + WC + c]); atomicAdd(gin + base + WC + c, w3 * tg); }
                            if (pt.flags & F_C4) { v4 = to_acc<T>(in[base + WC + p.C + c]); atomicAdd(gin + base + WC + p.C + c, w4 * tg); }
                            const A val = (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                            const A gw = pt.hh * (v2 - v1) + pt.lh * (v4 - v3);
                            const A gh = pt.hw * (v3 - v1) + pt.lw * (v4 - v2);
                            s_m += tgrad * val;
                            s_w += gw * tg;
                            s_h += gh * tg;
                        }
                    }
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) {
                        s_m += __shfl_xor_sync(0xffffffffu, s_m, o2);
                        s_w += __shfl_xor_sync(0xffffffffu, s_w, o2);
                        s_h += __shfl_xor_sync(0xffffffffu, s_h, o2);
                    }
                    if (lane == 0) { red[warp][0] = s_m; red[warp][1] = s_w; red[warp][2] = s_h; }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        gmsk[unit * p.P + k] = from_acc<T, A>(red[0][0] + red[1][0]);
                        goff[(unit * p.P + k) * 2] = from_acc<T, A>(scale * (red[0][1] + red[1][1]));
                        goff[(unit * p.P + k) * 2 + 1] = from_acc<T, A>(scale * (red[0][2] + red[1][2]));
                    }
                    __syncthreads();
                    ++k;
                }
    }
}

// fp32 accumulation image -> 16-bit grad_input (dcnv3_cuda.cu:168-173)
template <typename T>
__global__ void __launch_bounds__(256) cast_from_f32(const float *__restrict__ src, T *__restrict__ dst, long long n) {
    const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = *reinterpret_cast<const float4 *>(src + i), b = *reinterpret_cast<const float4 *>(src + i + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        Vec<T, 8>::store_stream(dst + i, v);
    } else {
        for (long long k = i; k < n; ++k) dst[k] = from_acc<T, float>(src[k]);
    }
}

// ---------------------------------------------------------------------------------------------------
// Index hook (parity of floor()/bounds): one thread per unit.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
dcnv3_index_kernel(const T *__restrict__ off, int *__restrict__ hw_low, unsigned char *__restrict__ flags,
                   const __grid_constant__ KParams p) {
    using A = typename AccOf<T>::type;
    const long long unit = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (unit >= p.n_units) return;
    const long long q = unit / p.G;
    const int ow = (int)(q % p.Wo), oh = (int)((q / p.Wo) % p.Ho);
    const A scale = (A)p.scale;
    const A p0_h_ = origin<A>(p.base_h + oh * p.sh, p.half_h, scale);
    const A p0_w_ = origin<A>(p.base_w + ow * p.sw, p.half_w, scale);
    const int ch = p.kh / 2, cw = p.kw / 2;
    long long k = unit * p.P;
    for (int i = 0; i < p.kw; ++i)
        for (int j = 0; j < p.kh; ++j)
            if (i != cw || j != ch || !p.remove_center) {
                Point<A> pt;
                locate<A>(pt, p0_h_, p0_w_, j * p.dh, i * p.dw, to_acc<T>(off[2 * k]), to_acc<T>(off[2 * k + 1]), scale,
                          p.H, p.W);
                hw_low[2 * k] = pt.h_low;
                hw_low[2 * k + 1] = pt.w_low;
                flags[k] = (unsigned char)pt.flags;
                ++k;
            }
}

}  // namespace gp
